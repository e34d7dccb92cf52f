#!/bin/bash
# round 2, call t: where the start-up of the `quack` process goes
OUT=gpurun_out/${1:-r02t}
mkdir -p $OUT
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
$G $D/s_1.fq.gz 7 1 0 200000 150 150 0.1 bgzf 1 8
AD=tests/golden/adapters_all.fa
for rep in 1 2 3 4; do
  echo "--- run $rep" >> $OUT/create_timing.txt
  QB_CREATE_TIMING=1 QB_STATS_JSON=$OUT/cli.json quack_b200/bin/quack -u $D/s_1.fq.gz -a $AD 2>> $OUT/create_timing.txt > /dev/null
  python -c "
import json; d=json.load(open('$OUT/cli.json')); print('create_s %.3f total %.3f' % (d['create_s'], d['total_s']))" >> $OUT/create_timing.txt
done
echo "--- lazy loading off" >> $OUT/create_timing.txt
CUDA_MODULE_LOADING=EAGER QB_CREATE_TIMING=1 quack_b200/bin/quack -u $D/s_1.fq.gz -a $AD 2>> $OUT/create_timing.txt > /dev/null
cat $OUT/create_timing.txt
nvidia-smi -q | grep -i "persistence" | head -2
rm -rf $D
