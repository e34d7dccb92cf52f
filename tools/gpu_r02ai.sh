#!/bin/bash
# round 2, call ai: compute-sanitizer memcheck over the whole parity suite (period / flat / fused / simple kernels)
OUT=gpurun_out/${1:-r02ai}
mkdir -p $OUT
( time timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_parity.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $OUT/pytest_parity.log 2>&1
echo "parity rc=$?" | tee -a $OUT/summary.txt
head -n 3 $OUT/pytest_parity.log | tee -a $OUT/summary.txt
tail -n 3 $OUT/memcheck_parity.log | tee -a $OUT/summary.txt
