#!/bin/bash
# round 2, call al: the default bench line with the final code; `quack` on 10 M pairs of BGZF, host reader vs device inflate
OUT=gpurun_out/${1:-r02al}
mkdir -p $OUT
( time python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err
python - <<PY
import json
b = json.load(open("$OUT/bench.json"))
print("value %.4e frac %.4f e2e %.4e" % (b["value"], b["roofline"]["frac"], b["e2e"]["value"]))
print("e2e_compressed", b.get("e2e_compressed"))
for k, v in b.get("e2e_file", {}).items():
    if isinstance(v, dict): print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("seconds", "create_s", "stream_s", "speedup_vs_reference", "svg_identical_to_reference")})
PY
D=/dev/shm/qbm; mkdir -p $D
G=quack_b200/bin/qb_gen_fastq
$G $D/b_1.fq.gz 7 1 0 10000000 150 150 0.1 bgzf 1 8 &
$G $D/b_2.fq.gz 7 2 0 10000000 150 150 0.1 bgzf 1 8 &
wait
AD=tests/golden/adapters_all.fa
for mode in host dev; do
for rep in 1 2 3; do
  if [ $mode = dev ]; then export QB_DEVICE_INFLATE=1; else export QB_DEVICE_INFLATE=0; fi
  QB_STATS_JSON=$OUT/cli.json quack_b200/bin/quack -1 $D/b_1.fq.gz -2 $D/b_2.fq.gz -a $AD -n x 2>> $OUT/cli.err > $OUT/cli_$mode.svg
  python -c "
import json; d=json.load(open('$OUT/cli.json')); print('$mode reads', d['reads'], 'create_s %.3f after_create %.3f total %.3f' % (d['create_s'], d['stream_s']-d['create_s'], d['total_s']))" | tee -a $OUT/cli_final.txt
done
done
cmp $OUT/cli_host.svg $OUT/cli_dev.svg && echo "svg identical" | tee -a $OUT/cli_final.txt
rm -f $OUT/*.svg; rm -rf $D
