#!/bin/bash
# round 2, call d: -a path with clean codes + strength-reduced addresses; tuning variants; full GPU suite
OUT=gpurun_out/${1:-r02d}
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py 16000000 >> $OUT/quick_bench.jsonl 2>&1
for v in "QB_PT_WARPS=16" "QB_PT_WARPS=16 QB_PT_BYTES=1800" "QB_PT_STAGES=2 QB_PT_WARPS=24"; do
  env $v QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py 16000000 >> $OUT/quick_bench_variants.jsonl 2>&1
done
for mode in ad; do
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:period_kernel -s 1 -c 1 \
    -o $OUT/period_${mode}_full -f python tools/profile_target.py $mode 2000000 150 150 3 > $OUT/ncu_full_$mode.log 2>&1
  ncu -i $OUT/period_${mode}_full.ncu-rep --page raw --csv > $OUT/period_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/period_${mode}_full.ncu-rep --page source --csv > $OUT/period_${mode}_full.src.csv 2>/dev/null
  rm -f $OUT/period_${mode}_full.ncu-rep
done
tail -5 $OUT/pytest_gpu.log; cat $OUT/quick_bench.jsonl $OUT/quick_bench_variants.jsonl
