#!/bin/bash
OUT=gpurun_out/${1:-r02aa}
mkdir -p $OUT
for cb in 0 3328 2752 2176; do
  echo "chunk $cb" >> $OUT/chunk_sweep.txt
  if [ $cb = 0 ]; then unset QB_FLAT_CHUNK; else export QB_FLAT_CHUNK=$cb; fi
  QB_QUICK_KERNELS=0 timeout 600 python tools/quick_bench.py 4000000 2>&1 | grep "35, 300" | cut -c1-200 >> $OUT/chunk_sweep.txt
done
cat $OUT/chunk_sweep.txt
