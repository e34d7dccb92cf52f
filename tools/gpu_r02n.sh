#!/bin/bash
# round 2, call n: text path parity tests; the `quack` program on 10 M pairs, host framing vs device framing
OUT=gpurun_out/${1:-r02n}
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -q ) > $OUT/pytest_text.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_text.log
tail -40 $OUT/pytest_text.log
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
N=${2:-10000000}
for fmt in gz bgzf; do
  $G $D/${fmt}_1.fq.gz 7 1 0 $N 150 150 0.1 $fmt 1 8 &
  $G $D/${fmt}_2.fq.gz 7 2 0 $N 150 150 0.1 $fmt 1 8 &
  wait
done
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
python - <<PY
import json
for l in open("$OUT/cli_framing.txt"):
    if "{" not in l: print(l.strip()); continue
    h, j = l.split("{", 1); d = json.loads("{" + j)
    print(h, "reads", d["reads"], "bases", d["bases"], "launches", d["launches"], "create_s %.3f stream_s %.3f after_create %.3f total %.3f" % (d["create_s"], d["stream_s"], d["stream_s"] - d["create_s"], d["total_s"]))
PY
cat $OUT/*.err | sort | uniq -c | head
rm -f $OUT/*.svg
rm -rf $D
