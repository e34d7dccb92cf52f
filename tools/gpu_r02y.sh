#!/bin/bash
OUT=gpurun_out/${1:-r02y}
mkdir -p $OUT
python tools/step_overhead.py 2>&1 | tee $OUT/step_overhead.txt
python tools/step_overhead.py 2>&1 | tee -a $OUT/step_overhead.txt
