#!/bin/bash
# round 2, call w (4 GPUs): the bench line at N=4 (value and e2e scaling, concurrent-H2D ceiling), multi-GPU tests
OUT=gpurun_out/${1:-r02w}
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/nproc.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 4 --steps 5 --warmup 3 ) > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
cat $OUT/bench_4gpu.json | cut -c1-2500; tail -3 $OUT/bench_4gpu.err
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_configs.py -m gpu -q ) > $OUT/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multi.log
tail -5 $OUT/pytest_multi.log
cat $OUT/nproc.txt
