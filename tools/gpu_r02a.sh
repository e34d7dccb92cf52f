#!/bin/bash
# round 2, call a: parity after the ring-ownership / growing-accumulator changes + baseline kernel numbers
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt; lscpu | head -30 >> $OUT/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
QB_QUICK_KERNELS=0 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
tail -5 $OUT/pytest_gpu.log; cat $OUT/quick_bench.jsonl
