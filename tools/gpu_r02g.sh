#!/bin/bash
# round 2, call g: flat kernel v2 (lane <-> word), member pool; full suite; decode + CLI benchmarks
OUT=gpurun_out/${1:-r02g}
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flat" ) > $OUT/pytest_flat.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_flat.log
( time timeout 1800 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
QB_QUICK_KERNELS=0,2 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=5 QB_QUICK_LENS=50,100,150,151,250,300 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench_flat_uniform.jsonl 2>&1
for mode in ad noad; do
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_kernel -s 1 -c 1 \
    -o $OUT/flat_${mode}_full -f python tools/profile_target.py $mode 2000000 35 300 3 > $OUT/ncu_full_$mode.log 2>&1
  ncu -i $OUT/flat_${mode}_full.ncu-rep --page raw --csv > $OUT/flat_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/flat_${mode}_full.ncu-rep --page source --csv > $OUT/flat_${mode}_full.src.csv 2>/dev/null
  rm -f $OUT/flat_${mode}_full.ncu-rep
done
( time timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench.json 2> $OUT/bench.err
tail -5 $OUT/pytest_flat.log; tail -5 $OUT/pytest_gpu.log; cat $OUT/quick_bench.jsonl
