#!/bin/bash
OUT=gpurun_out/${1:-r02ad}
mkdir -p $OUT
python bench.py --no-e2e-file --no-other-kernels --no-cpu-baseline > $OUT/bench_quick.json 2> $OUT/bench_quick.err
tail -3 $OUT/bench_quick.err
python - <<PY
import json
b = json.load(open("$OUT/bench_quick.json"))
print("value %.4e ms/step %.3f frac %.3f" % (b["value"], b["ms_per_step"], b["roofline"]["frac"]))
print(json.dumps(b["e2e"], indent=0)); print(b["clocks"])
PY
