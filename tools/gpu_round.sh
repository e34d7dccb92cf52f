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
timeout 600 python tools/decode_bench.py 2000000 16 > $OUT/decode_bench.jsonl 2>&1
timeout 900 python tools/cli_bench.py 2000000 1,4,8 > $OUT/cli_bench.jsonl 2>&1
QB_QUICK_KERNELS=0,4,2,3 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=4,2 QB_QUICK_LENS=50,76,100,126,200,256 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench_lens.jsonl 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launch list of the same bench command (cold-cache, serialised): share of the step per kernel
QB_BENCH_E2E_STEPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# one full capture of the dominant kernel (period kernel, the AUTO choice for uniform-length batches)
for mode in ad noad; do
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:period_kernel -s 1 -c 1 \
    -o $OUT/period_${mode}_full -f python tools/profile_target.py $mode 2000000 150 150 3 > $OUT/ncu_full_$mode.log 2>&1
  ncu -i $OUT/period_${mode}_full.ncu-rep --page raw --csv > $OUT/period_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/period_${mode}_full.ncu-rep --page source --csv > $OUT/period_${mode}_full.src.csv 2>/dev/null
done
ls -la $OUT
