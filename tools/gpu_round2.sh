#!/bin/bash
# Round-2 evidence run (1 x B200): GPU parity suite, bench (both arms), ncu launch list of the bench command, full ncu
# captures of the period kernel (-a / no adapters, at the bench's 10 M reads per launch for the DRAM traffic) and of
# the flat kernel, kernel sweeps, host decode and CLI benchmarks.   usage: tools/gpu_round2.sh [tag]
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
QB_BENCH_E2E_STEPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-file --no-other-kernels > $OUT/bench_under_ncu.log 2>&1
QB_QUICK_KERNELS=0,2 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py 16000000 >> $OUT/quick_bench.jsonl 2>&1
QB_QUICK_KERNELS=4 QB_QUICK_LENS=50,76,100,126,151,200,256 timeout 600 python tools/quick_bench.py 4000000 > $OUT/quick_bench_lens.jsonl 2>&1
timeout 600 python tools/decode_bench.py 2000000 16 > $OUT/decode_bench.jsonl 2>&1
for lvl in 1 6; do timeout 300 python tools/inflate_bench.py 4000000 $lvl >> $OUT/inflate_bench.jsonl 2>> $OUT/inflate_bench.err; done
timeout 300 python tools/inflate_bench.py 100000 1 >> $OUT/inflate_bench.jsonl 2>> $OUT/inflate_bench.err
for mode in ad noad; do
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:period_kernel -s 1 -c 1 \
    -o $OUT/period_${mode}_full -f python tools/profile_target.py $mode 10000000 150 150 3 > $OUT/ncu_period_$mode.log 2>&1
  ncu -i $OUT/period_${mode}_full.ncu-rep --page raw --csv > $OUT/period_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/period_${mode}_full.ncu-rep --page source --csv > $OUT/period_${mode}_full.src.csv 2>/dev/null
  QB_PROFILE_KERNEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_kernel -s 1 -c 1 \
    -o $OUT/flat_${mode}_full -f python tools/profile_target.py $mode 5000000 35 300 3 > $OUT/ncu_flat_$mode.log 2>&1
  ncu -i $OUT/flat_${mode}_full.ncu-rep --page raw --csv > $OUT/flat_${mode}_full.raw.csv 2>/dev/null
  ncu -i $OUT/flat_${mode}_full.ncu-rep --page source --csv > $OUT/flat_${mode}_full.src.csv 2>/dev/null
  rm -f $OUT/period_${mode}_full.ncu-rep $OUT/flat_${mode}_full.ncu-rep
done
ls -la $OUT
