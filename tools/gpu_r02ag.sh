#!/bin/bash
# round 2, call ag: compute-sanitizer memcheck over the new kernels (flat after the load changes, framing, inflate, extras)
OUT=gpurun_out/${1:-r02ag}
mkdir -p $OUT
run() { # name, pytest args...
  local name=$1; shift
  ( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_$name.log python -m pytest "$@" -m gpu -q -x ) > $OUT/pytest_$name.log 2>&1
  echo "$name rc=$?" | tee -a $OUT/summary.txt
  grep -c "Invalid\|out of bounds\|misaligned" $OUT/memcheck_$name.log | tee -a $OUT/summary.txt
  tail -3 $OUT/memcheck_$name.log | tee -a $OUT/summary.txt
}
run extras tests/test_gpu_extras.py -k "equal_the_cpu or text_path"
run inflate tests/test_gpu_inflate.py -k "fastq or runs or tiny or empty or refuses or damaged or bgzf_submit_equals"
run text tests/test_gpu_text.py -k "cut_anywhere or small_chunks or truncated or refused or two_mates"
run flat tests/test_gpu_parity.py -k "flat"
