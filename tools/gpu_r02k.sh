#!/bin/bash
# round 2, call k (2 GPUs): multi-GPU parity (in-process reduce, one process per GPU, config 3 sharded) and the
# 2-GPU bench line with the concurrent-H2D ceiling
OUT=gpurun_out/${1:-r02k}
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( time timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_configs.py -m gpu -q ) > $OUT/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 ) > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -5 $OUT/pytest_multi.log; cat $OUT/bench_2gpu.json | cut -c1-3000; tail -3 $OUT/bench_2gpu.err
