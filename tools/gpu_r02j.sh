#!/bin/bash
# round 2, call j: full GPU suite after the flat -a fix and the wtile removal; bench; decode / CLI benchmarks
OUT=gpurun_out/${1:-r02j}
mkdir -p $OUT
( time timeout 1800 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
QB_QUICK_KERNELS=0 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
timeout 600 python tools/decode_bench.py 2000000 16 > $OUT/decode_bench.jsonl 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
tail -5 $OUT/pytest_gpu.log; cat $OUT/quick_bench.jsonl; cat $OUT/decode_bench.jsonl
