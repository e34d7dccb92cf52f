#!/bin/bash
# round 2, call h: flat kernel v3 (straight-line steps, two register groups): parity, -a debug, speed, ncu
OUT=gpurun_out/${1:-r02h}
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "flat" ) > $OUT/pytest_flat.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_flat.log
timeout 600 python tools/debug_flat.py > $OUT/debug_flat.log 2>&1
QB_QUICK_KERNELS=0 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
for mode in noad; do
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_kernel -s 1 -c 1 \
    -o $OUT/flat_${mode}_full -f python tools/profile_target.py $mode 2000000 35 300 3 > $OUT/ncu_full_$mode.log 2>&1
  ncu -i $OUT/flat_${mode}_full.ncu-rep --page raw --csv > $OUT/flat_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/flat_${mode}_full.ncu-rep --page source --csv > $OUT/flat_${mode}_full.src.csv 2>/dev/null
  rm -f $OUT/flat_${mode}_full.ncu-rep
done
tail -5 $OUT/pytest_flat.log; cat $OUT/debug_flat.log | tail -20; cat $OUT/quick_bench.jsonl
