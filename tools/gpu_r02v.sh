#!/bin/bash
# round 2, call v: flat kernel after the load / lookup trimming: parity, ragged kernel-only numbers
OUT=gpurun_out/${1:-r02v}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "flat or ragged or config4 or Flat" ) > $OUT/pytest_flat.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_flat.log
tail -5 $OUT/pytest_flat.log
QB_QUICK_KERNELS=0 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
cat $OUT/quick_bench.jsonl | cut -c1-250
