#!/bin/bash
# GPU iteration on the period kernel (v5): parity tests, kernel-only throughput, stage / tile-size sweep,
# full ncu capture with and without adapters.
# usage: tools/gpu_period.sh <tag> [pytest -k expression] [ncu: 0/1]
TAG=${1:-p}
KEXPR=${2:-period}
NCU=${3:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$KEXPR" ) > $OUT/pytest_parity.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_parity.log
tail -15 $OUT/pytest_parity.log
QB_QUICK_KERNELS=4,2 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
cat $OUT/quick_bench.jsonl
QB_QUICK_KERNELS=4,2 QB_QUICK_LENS=${LENS:-50,76,100,126,200,256} timeout 300 python tools/quick_bench.py 4000000 >> $OUT/sweep.jsonl 2>&1
QB_PT_NATURAL=1 QB_QUICK_KERNELS=4 QB_QUICK_ONLY150=1 timeout 300 python tools/quick_bench.py 4000000 >> $OUT/sweep.jsonl 2>&1
for w in 16 20 24; do for st in 2 3; do
  QB_PT_WARPS=$w QB_PT_STAGES=$st QB_QUICK_KERNELS=4 QB_QUICK_ONLY150=1 timeout 300 python tools/quick_bench.py 4000000 2>&1 | grep -v "^Traceback\|^  File\|^    " >> $OUT/sweep.jsonl
done; done
cat $OUT/sweep.jsonl
if [ "$NCU" = "1" ]; then
for mode in ad noad; do
  QB_PROFILE_KERNEL=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'period_kernel' -s 1 -c 1 \
    -o $OUT/${mode}_full -f python tools/profile_target.py $mode 2000000 150 150 3 > $OUT/ncu_$mode.log 2>&1
  tail -1 $OUT/ncu_$mode.log
  ncu -i $OUT/${mode}_full.ncu-rep --page raw --csv > $OUT/${mode}_raw.csv 2>/dev/null
done
fi
