#!/bin/bash
# round 2, call e: tile staging A/B -- per-lane cp.async (default build) against 1-D TMA bulk copies (-DQB_PT_TMA=1)
OUT=gpurun_out/${1:-r02e}
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "period" ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
for lib in "" "quack_b200/lib/libquack_b200_tma.so"; do
  for n in 4000000 16000000; do
    QB_LIB=${lib:+$PWD/$lib} QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 600 python tools/quick_bench.py $n >> $OUT/quick_bench.jsonl 2>&1
  done
  QB_LIB=${lib:+$PWD/$lib} QB_QUICK_KERNELS=4 QB_QUICK_LENS=100,126,200,256 timeout 600 python tools/quick_bench.py 4000000 >> $OUT/quick_bench_lens.jsonl 2>&1
done
tail -3 $OUT/pytest_gpu.log; cat $OUT/quick_bench.jsonl
