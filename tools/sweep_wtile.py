#!/usr/bin/env python
"""Kernel-only timing of the warp-tile kernel on 150-bp reads for several reads-per-tile values
(QB_WT_READS) -- run once per build (QB_LIB = a library compiled with another QB_WW).
usage: sweep_wtile.py [n_reads] [R,R,...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import qb_testutil as util
from quack_b200 import capi

HBM = 6545.3
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
keys = np.concatenate([capi.adapter_record_keys(r) for r in util.adapter_records()])
for ad in (None, keys):
    with capi.Context(150, adapter_keys=ad, kernel=capi.KERNEL_WTILE) as ctx:
        b = ctx.generate(2, 1, 0, n, 150, 150, 0.1)
        nr, nb = b.info
        for r in rs:
            if r:
                os.environ["QB_WT_READS"] = str(r)
            else:
                os.environ.pop("QB_WT_READS", None)
            try:
                avg, mn = b.time(0, warmup=2, iters=5, flush_l2=False)
            except capi.QbError as e:
                print(json.dumps({"R": r, "adapters": ad is not None, "error": str(e)[:80]}), flush=True)
                continue
            alg = 2 * nb + 8 * nr
            print(json.dumps({"lib": os.path.basename(os.environ.get("QB_LIB", "default")), "R": r, "adapters": ad is not None,
                              "ms": round(avg, 4), "GBps": round(alg / avg / 1e6, 1),
                              "frac_hbm": round(alg / avg / 1e6 / HBM, 4)}), flush=True)
        b.free()
