#!/bin/bash
# round 2, call p: inflate kernel throughput; `quack` on 10 M pairs of BGZF: host reader vs device inflate
OUT=gpurun_out/${1:-r02p}
mkdir -p $OUT
for lvl in 1 6; do python tools/inflate_bench.py 4000000 $lvl >> $OUT/inflate_bench.jsonl 2>> $OUT/inflate_bench.err; done
cat $OUT/inflate_bench.jsonl; tail -3 $OUT/inflate_bench.err
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
N=${2:-10000000}
$G $D/b_1.fq.gz 7 1 0 $N 150 150 0.1 bgzf 1 8 &
$G $D/b_2.fq.gz 7 2 0 $N 150 150 0.1 bgzf 1 8 &
wait
AD=tests/golden/adapters_all.fa
for mode in host dev dev64; do
  for rep in 1 2 3; do
    case $mode in
      host) E="QB_DEVICE_INFLATE=0" ;;
      dev) E="QB_DEVICE_INFLATE=1" ;;
      dev64) E="QB_DEVICE_INFLATE=1 QB_BATCH_MB=64" ;;
    esac
    env $E QB_VERBOSE=1 QB_STATS_JSON=$OUT/cli_$mode.json quack_b200/bin/quack -1 $D/b_1.fq.gz -2 $D/b_2.fq.gz -a $AD -n x > $OUT/cli_$mode.svg 2> $OUT/cli_$mode.err
    echo "bgzf mode=$mode rep=$rep rc=$? $(cat $OUT/cli_$mode.json)" >> $OUT/cli_inflate.txt
  done
done
cmp $OUT/cli_host.svg $OUT/cli_dev.svg && cmp $OUT/cli_host.svg $OUT/cli_dev64.svg && echo "svg identical" >> $OUT/cli_inflate.txt
python - <<PY
import json
for l in open("$OUT/cli_inflate.txt"):
    if "{" not in l: print(l.strip()); continue
    h, j = l.split("{", 1); d = json.loads("{" + j)
    print(h, "reads", d["reads"], "launches", d["launches"], "create_s %.3f stream_s %.3f after_create %.3f total %.3f" % (d["create_s"], d["stream_s"], d["stream_s"] - d["create_s"], d["total_s"]))
PY
cat $OUT/*.err | sort | uniq -c | head
rm -f $OUT/*.svg
rm -rf $D
