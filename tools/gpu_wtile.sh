#!/bin/bash
# GPU iteration on the warp-tile kernel: parity tests, kernel sweep over warps per CTA (alternate builds
# lib/libqb_w<N>.so) and reads per tile, full ncu capture with and without adapters.
# usage: tools/gpu_wtile.sh <tag> [R list for the sweep]
TAG=${1:-w}
RS=${2:-0,4,6,8,9,10,12,13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $OUT/pytest_parity.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_parity.log
tail -5 $OUT/pytest_parity.log
QB_QUICK_KERNELS=3,2 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
cat $OUT/quick_bench.jsonl
timeout 300 python tools/sweep_wtile.py 4000000 $RS > $OUT/sweep.jsonl 2>&1
for lib in quack_b200/lib/libqb_w*.so; do
  QB_LIB=$PWD/$lib timeout 300 python tools/sweep_wtile.py 4000000 $RS >> $OUT/sweep.jsonl 2>&1
done
cat $OUT/sweep.jsonl
for mode in ad noad; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wtile_kernel' -s 1 -c 1 \
    -o $OUT/${mode}_full -f python tools/profile_target.py $mode 2000000 150 150 3 > $OUT/ncu_$mode.log 2>&1
  tail -1 $OUT/ncu_$mode.log
done
