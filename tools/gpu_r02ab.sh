#!/bin/bash
OUT=gpurun_out/${1:-r02ab}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_extras.py -m gpu -q -x ) > $OUT/pytest_extras.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_extras.log
tail -40 $OUT/pytest_extras.log
