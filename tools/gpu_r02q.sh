#!/bin/bash
# round 2, call q: where the device-inflate CLI spends its time (host stage timers, kernel launch list)
OUT=gpurun_out/${1:-r02q}
mkdir -p $OUT
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
$G $D/b_1.fq.gz 7 1 0 10000000 150 150 0.1 bgzf 1 8 &
$G $D/b_2.fq.gz 7 2 0 10000000 150 150 0.1 bgzf 1 8 &
$G $D/s_1.fq.gz 7 1 0 1000000 150 150 0.1 bgzf 1 8 &
wait
AD=tests/golden/adapters_all.fa
for mb in 16 64; do
for rep in 1 2; do
  QB_DEVICE_INFLATE=1 QB_BATCH_MB=$mb QB_VERBOSE=2 QB_STATS_JSON=$OUT/cli.json quack_b200/bin/quack -1 $D/b_1.fq.gz -2 $D/b_2.fq.gz -a $AD -n x 2>> $OUT/cli_timers.txt > /dev/null
  echo "mb=$mb $(cat $OUT/cli.json)" >> $OUT/cli_timers.txt
done
done
cat $OUT/cli_timers.txt | cut -c1-400
QB_DEVICE_INFLATE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_cli.csv quack_b200/bin/quack -u $D/s_1.fq.gz -a $AD > /dev/null 2> $OUT/ncu.err
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/launches_cli.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    agg[r[ki][:60]][0] += 1; agg[r[ki][:60]][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-62s n=%4d  %10.1f us  %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
PY
rm -rf $D
