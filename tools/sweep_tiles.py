#!/usr/bin/env python
"""Kernel-only timing of the fused kernel on 150-bp reads for several tile geometries
(QB_TILE_PASSES x QB_STAGES), with and without the adapter set."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, qb_testutil as util
from quack_b200 import capi
keys = np.concatenate([capi.adapter_record_keys(r) for r in util.adapter_records()])
out = {}
for name, ad, lmin, lmax, cap in (("noad", None, 150, 150, 150), ("ad", keys, 150, 150, 150), ("rag_ad", keys, 35, 300, 304)):
    with capi.Context(cap, adapter_keys=ad, kernel=capi.KERNEL_FUSED) as ctx:
        b = ctx.generate(2, 1, 0, 6000000 if lmin == 150 else 3000000, lmin, lmax, 0.1)
        nr, nb = b.info
        avg, mn = b.time(0, warmup=2, iters=5, flush_l2=False)
        out[name] = round((2 * nb + 8 * nr) / avg / 1e6 / 6548.5, 4)
        b.free()
print(json.dumps(out))
''' % (ROOT, ROOT)

for passes in sys.argv[1].split(","):
    for stages in sys.argv[2].split(","):
        env = dict(os.environ, QB_TILE_PASSES=passes, QB_STAGES=stages)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(f"passes={passes} stages={stages} lib={os.path.basename(os.environ.get('QB_LIB', 'default'))}",
              r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
