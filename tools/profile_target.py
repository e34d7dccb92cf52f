#!/usr/bin/env python
"""Small fixed workload for ncu: a few launches of the fused kernel on generated reads.
usage: profile_target.py [ad|noad] [n_reads] [len_min] [len_max] [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import qb_testutil as util
from quack_b200 import capi

ad = (sys.argv[1] if len(sys.argv) > 1 else "ad") == "ad"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
lmin = int(sys.argv[3]) if len(sys.argv) > 3 else 150
lmax = int(sys.argv[4]) if len(sys.argv) > 4 else 150
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 3
keys = np.concatenate([capi.adapter_record_keys(r) for r in util.adapter_records()]) if ad else None
with capi.Context(max(lmax, 11), adapter_keys=keys, kernel=int(os.environ.get("QB_PROFILE_KERNEL", capi.KERNEL_AUTO))) as ctx:
    b = ctx.generate(2, 1, 0, n, lmin, lmax, 0.1)
    for _ in range(launches):
        b.run(0)
    r = ctx.finish(0)
    print("reads", r.n_reads, "max_length", r.max_length)
    b.free()
