#!/bin/bash
# GPU iteration on the warp-tile kernel: parity tests, kernel sweep over warps per CTA (alternate builds
# lib/libqb_w<N>.so) and reads per tile, full ncu capture with and without adapters.
# usage: tools/gpu_wtile.sh <tag>
TAG=${1:-w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $OUT/pytest_parity.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_parity.log
tail -5 $OUT/pytest_parity.log
QB_QUICK_KERNELS=3,2 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
for lib in quack_b200/lib/libqb_w*.so; do
  QB_LIB=$PWD/$lib QB_QUICK_KERNELS=3 timeout 300 python tools/quick_bench.py 4000000 >> $OUT/quick_bench_alt.jsonl 2>&1
done
cat $OUT/quick_bench.jsonl $OUT/quick_bench_alt.jsonl | grep -v simple
