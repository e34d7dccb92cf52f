#!/bin/bash
# round 2, call ah: compute-sanitizer racecheck (shared-memory hazards) over the inflate, framing and flat kernels
OUT=gpurun_out/${1:-r02ah}
mkdir -p $OUT
run() {
  local name=$1; shift
  ( time timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --log-file $OUT/racecheck_$name.log python -m pytest "$@" -m gpu -q -x ) > $OUT/pytest_$name.log 2>&1
  echo "$name rc=$?" | tee -a $OUT/summary.txt
  grep -i "hazard\|ERROR SUMMARY" $OUT/racecheck_$name.log | sort | uniq -c | head -8 | tee -a $OUT/summary.txt
  head -n 3 $OUT/pytest_$name.log | tee -a $OUT/summary.txt
}
run inflate tests/test_gpu_inflate.py -k "fastq or runs or skewed"
run text tests/test_gpu_text.py -k "cut_anywhere and 150"
run flat tests/test_gpu_parity.py -k "flat and (35 or 16)"
