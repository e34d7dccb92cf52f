#!/bin/bash
# round 2, call b: new -a path of the period kernel (persistent hit queue, packed-code ring): parity, speed, ncu
OUT=gpurun_out/${1:-r02b}
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ring.py -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py 16000000 >> $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=4 QB_QUICK_LENS=50,76,100,126,151,200,256 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench_lens.jsonl 2>&1
for mode in ad noad; do
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:period_kernel -s 1 -c 1 \
    -o $OUT/period_${mode}_full -f python tools/profile_target.py $mode 2000000 150 150 3 > $OUT/ncu_full_$mode.log 2>&1
  ncu -i $OUT/period_${mode}_full.ncu-rep --page raw --csv > $OUT/period_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/period_${mode}_full.ncu-rep --page source --csv > $OUT/period_${mode}_full.src.csv 2>/dev/null
  rm -f $OUT/period_${mode}_full.ncu-rep
done
tail -5 $OUT/pytest_gpu.log; cat $OUT/quick_bench.jsonl
