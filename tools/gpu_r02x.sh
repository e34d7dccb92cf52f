#!/bin/bash
# round 2, call x: the default bench line and the reference arm with everything of this round in; new CLI tests
OUT=gpurun_out/${1:-r02x}
mkdir -p $OUT
( time python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err
python - <<PY
import json
b = json.load(open("$OUT/bench.json"))
print("value %.3e frac %.3f e2e %.3e ms/step %.3f" % (b["value"], b["roofline"]["frac"], b["e2e"]["value"], b["ms_per_step"]))
print(json.dumps(b.get("roofline_other_kernels"), indent=0)[:1500])
for k, v in b.get("e2e_file", {}).items():
    if isinstance(v, dict): print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("seconds", "value", "create_s", "stream_s", "speedup_vs_reference", "svg_identical_to_reference")})
    else: print(k, v)
PY
( time timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -q -k tricky ) > $OUT/pytest_tricky.log 2>&1
tail -4 $OUT/pytest_tricky.log
