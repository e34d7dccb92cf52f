#!/bin/bash
OUT=gpurun_out/${1:-r02i}
mkdir -p $OUT
timeout 900 python tools/debug_flat.py > $OUT/debug_flat.log 2>&1
timeout 900 python tools/debug_flat.py 26017 30011 17 17 > $OUT/debug_flat17.log 2>&1
tail -12 $OUT/debug_flat.log; tail -12 $OUT/debug_flat17.log
