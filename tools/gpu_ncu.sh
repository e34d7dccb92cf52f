#!/bin/bash
# Full ncu capture of the statistics kernel on 150-bp reads, with and without adapters.
# usage: tools/gpu_ncu.sh <tag> [kernel id: 3 wtile (default), 2 fused]
TAG=${1:-n}
export QB_PROFILE_KERNEL=${2:-3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for mode in ad noad; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wtile_kernel|fused_kernel' -s 1 -c 1 \
    -o $OUT/${mode}_full -f python tools/profile_target.py $mode 2000000 150 150 3 > $OUT/ncu_$mode.log 2>&1
  tail -2 $OUT/ncu_$mode.log
done
ls -la $OUT
