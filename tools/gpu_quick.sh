#!/bin/bash
# Short GPU iteration: parity tests, kernel sweep, full ncu capture of the fused kernel (with/without adapters).
# usage: tools/gpu_quick.sh <tag> [alt-lib.so]
TAG=${1:-q}
ALT=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 \
  -o $OUT/fused_ad_full -f python tools/profile_target.py ad 2000000 150 150 3 > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 \
  -o $OUT/fused_noad_full -f python tools/profile_target.py noad 2000000 150 150 3 > $OUT/ncu_full_noad.log 2>&1
if [ -n "$ALT" ]; then
  export QB_LIB=$PWD/$ALT
  ( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $OUT/pytest_gpu_alt.log 2>&1
  timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench_alt.jsonl 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 \
    -o $OUT/fused_ad_full_alt -f python tools/profile_target.py ad 2000000 150 150 3 > $OUT/ncu_full_alt.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 \
    -o $OUT/fused_noad_full_alt -f python tools/profile_target.py noad 2000000 150 150 3 > $OUT/ncu_full_noad_alt.log 2>&1
  echo "== alt"; tail -3 $OUT/pytest_gpu_alt.log; grep fused $OUT/quick_bench_alt.jsonl
fi
echo "== default"; tail -3 $OUT/pytest_gpu.log; grep fused $OUT/quick_bench.jsonl
