#!/bin/bash
OUT=gpurun_out/${1:-r02ak}
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_bgzf -s 2 -c 1 -o $OUT/inflate_ring -f python tools/inflate_bench.py 100000 1 > $OUT/ncu_inflate.log 2>&1
tail -2 $OUT/ncu_inflate.log
