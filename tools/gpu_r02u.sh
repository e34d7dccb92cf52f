#!/bin/bash
# round 2, call u: the whole -m gpu suite + smoke after the text / inflate additions
OUT=gpurun_out/${1:-r02u}
mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
