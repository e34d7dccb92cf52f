#!/bin/bash
# One gpurun call: GPU parity tests, microbench, quick kernel sweep, bench line, ncu launch list + full capture.
# usage: tools/gpu_round.sh [tag]   (outputs under gpurun_out/<tag>/)
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python tools/microbench.py > $OUT/microbench.txt 2>&1
timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launch list of the same bench command (cold-cache, serialised): share of the step per kernel
QB_BENCH_E2E_STEPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# one full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 \
  -o $OUT/fused_ad_full -f python tools/profile_target.py ad 2000000 150 150 3 > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 \
  -o $OUT/fused_noad_full -f python tools/profile_target.py noad 2000000 150 150 3 > $OUT/ncu_full_noad.log 2>&1
ls -la $OUT
