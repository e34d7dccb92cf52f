#!/bin/bash
# round 2, call r: inflate kernel after the latency work: tests, kernel bench (big and chunk-sized launches), CLI
OUT=gpurun_out/${1:-r02r}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_inflate.py -m gpu -q ) > $OUT/pytest_inflate.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_inflate.log
tail -5 $OUT/pytest_inflate.log
for n in 4000000 100000 400000; do python tools/inflate_bench.py $n 1 >> $OUT/inflate_bench.jsonl 2>> $OUT/inflate_bench.err; done
cat $OUT/inflate_bench.jsonl; tail -3 $OUT/inflate_bench.err
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
$G $D/b_1.fq.gz 7 1 0 10000000 150 150 0.1 bgzf 1 8 &
$G $D/b_2.fq.gz 7 2 0 10000000 150 150 0.1 bgzf 1 8 &
wait
AD=tests/golden/adapters_all.fa
for mb in 16 64; do
for rep in 1 2; do
  QB_DEVICE_INFLATE=1 QB_BATCH_MB=$mb QB_VERBOSE=2 QB_STATS_JSON=$OUT/cli.json quack_b200/bin/quack -1 $D/b_1.fq.gz -2 $D/b_2.fq.gz -a $AD -n x 2>> $OUT/cli_timers.txt > /dev/null
  python -c "
import json; d=json.load(open('$OUT/cli.json')); print('mb=$mb reads', d['reads'], 'create_s %.3f after_create %.3f total %.3f' % (d['create_s'], d['stream_s']-d['create_s'], d['total_s']))" >> $OUT/cli_timers.txt
done
done
cat $OUT/cli_timers.txt | cut -c1-400
rm -rf $D
