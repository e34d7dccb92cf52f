#!/bin/bash
# round 2, call m: the device-side framing path (qb_text_submit): parity tests, then the `quack` program on 2 M pairs
# (gzip members and BGZF) with the host reader framing vs the device framing
OUT=gpurun_out/${1:-r02m}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -q -x ) > $OUT/pytest_text.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_text.log
tail -30 $OUT/pytest_text.log
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
for fmt in gz bgzf; do
  $G $D/${fmt}_1.fq.gz 7 1 0 2000000 150 150 0.1 $fmt 1 &
  $G $D/${fmt}_2.fq.gz 7 2 0 2000000 150 150 0.1 $fmt 1 &
done
$G $D/rag_1.fq.gz 7 1 0 2000000 35 300 0.1 gz 1 &
wait
AD=tests/golden/adapters_all.fa
for fmt in gz bgzf; do
  for mode in 0 1; do
    for rep in 1 2 3; do
      QB_DEVICE_INFLATE=0 QB_DEVICE_FRAMING=$mode QB_VERBOSE=1 QB_STATS_JSON=$OUT/cli_${fmt}_$mode.json quack_b200/bin/quack -1 $D/${fmt}_1.fq.gz -2 $D/${fmt}_2.fq.gz -a $AD -n x > $OUT/cli_${fmt}_$mode.svg 2> $OUT/cli_${fmt}_$mode.err
      echo "$fmt framing=$mode rep=$rep rc=$? $(cat $OUT/cli_${fmt}_$mode.json)" >> $OUT/cli_framing.txt
    done
  done
  cmp $OUT/cli_${fmt}_0.svg $OUT/cli_${fmt}_1.svg && echo "$fmt svg identical" >> $OUT/cli_framing.txt
done
for mode in 0 1; do
  QB_DEVICE_INFLATE=0 QB_DEVICE_FRAMING=$mode QB_VERBOSE=1 QB_STATS_JSON=$OUT/cli_rag_$mode.json quack_b200/bin/quack -u $D/rag_1.fq.gz -a $AD > $OUT/cli_rag_$mode.svg 2> $OUT/cli_rag_$mode.err
  echo "ragged framing=$mode rc=$? $(cat $OUT/cli_rag_$mode.json)" >> $OUT/cli_framing.txt
done
cmp $OUT/cli_rag_0.svg $OUT/cli_rag_1.svg && echo "ragged svg identical" >> $OUT/cli_framing.txt
cat $OUT/cli_framing.txt | cut -c1-420
cat $OUT/*.err | sort | uniq -c | head
rm -f $OUT/*.svg
rm -rf $D
