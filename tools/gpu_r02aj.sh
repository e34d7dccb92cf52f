#!/bin/bash
OUT=gpurun_out/${1:-r02aj}
mkdir -p $OUT
for w in 0 24 16; do
  if [ $w = 0 ]; then unset QB_PT_WARPS; else export QB_PT_WARPS=$w; fi
  echo "QB_PT_WARPS=$w" >> $OUT/warps.txt
  QB_QUICK_KERNELS=0 QB_QUICK_ONLY150=1 timeout 300 python tools/quick_bench.py 10000000 2>&1 | grep '"len"' | cut -c1-200 >> $OUT/warps.txt
  python - >> $OUT/warps.txt <<PY
from quack_b200 import capi
import ctypes as C
out=(C.c_uint32*16)()
for ad in (0,1):
    rc=capi.lib().qb_period_plan_info(150, ad, out); print("plan ad=%d rc=%d"%(ad,rc), list(out))
PY
done
cat $OUT/warps.txt
