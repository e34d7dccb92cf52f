#!/bin/bash
# round 2, call o: device inflate + device framing tests
OUT=gpurun_out/${1:-r02o}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_text.py -m gpu -q ) > $OUT/pytest_text.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_text.log
tail -60 $OUT/pytest_text.log
