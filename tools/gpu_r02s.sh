#!/bin/bash
# round 2, call s: ncu capture of the inflate kernel (2000 blocks), ring-depth sweep of the device-inflate CLI
OUT=gpurun_out/${1:-r02s}
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_bgzf -s 2 -c 1 -o $OUT/inflate_full -f python tools/inflate_bench.py 400000 1 > $OUT/ncu_inflate.log 2>&1
tail -2 $OUT/ncu_inflate.log
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
$G $D/b_1.fq.gz 7 1 0 10000000 150 150 0.1 bgzf 1 8 &
$G $D/b_2.fq.gz 7 2 0 10000000 150 150 0.1 bgzf 1 8 &
wait
AD=tests/golden/adapters_all.fa
for cfg in "16 3" "16 6" "32 6" "64 6" "32 8"; do
set -- $cfg
for rep in 1 2; do
  QB_DEVICE_INFLATE=1 QB_BATCH_MB=$1 QB_RING=$2 QB_VERBOSE=2 QB_STATS_JSON=$OUT/cli.json quack_b200/bin/quack -1 $D/b_1.fq.gz -2 $D/b_2.fq.gz -a $AD -n x 2>> $OUT/cli_timers.txt > /dev/null
  python -c "
import json; d=json.load(open('$OUT/cli.json')); print('mb=$1 ring=$2 reads', d['reads'], 'create_s %.3f after_create %.3f total %.3f' % (d['create_s'], d['stream_s']-d['create_s'], d['total_s']))" >> $OUT/cli_timers.txt
done
done
grep -v "^quack" $OUT/cli_timers.txt
rm -rf $D
