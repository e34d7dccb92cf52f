#!/bin/bash
OUT=gpurun_out/${1:-r02ae}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_inflate.py -m gpu -q -x ) > $OUT/pytest_inflate.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_inflate.log
tail -5 $OUT/pytest_inflate.log
for n in 4000000 100000; do python tools/inflate_bench.py $n 1 >> $OUT/inflate_bench.jsonl 2>> $OUT/inflate_bench.err; done
python tools/inflate_bench.py 4000000 6 >> $OUT/inflate_bench.jsonl 2>> $OUT/inflate_bench.err
cat $OUT/inflate_bench.jsonl; tail -3 $OUT/inflate_bench.err
