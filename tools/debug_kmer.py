#!/usr/bin/env python
"""Debug aid: find the reads on which the fused kernel's adapter histogram differs from the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import capi

table = util.oracle_table()
keys = table.keys()
seq, qual, off, lens = util.random_batch(11, 20000, 150, 150, plant=0.2)


def sub(a, b):
    o0, o1 = int(off[a]), int(off[b - 1] + lens[b - 1])
    return seq[o0:o1].copy(), qual[o0:o1].copy(), (off[a:b] - off[a]).astype(np.uint32), lens[a:b].copy()


def gpu(batch):
    with capi.Context(150, adapter_keys=keys, kernel=capi.KERNEL_FUSED) as ctx:
        b = ctx.upload(*batch, max_len=150)
        b.run(0)
        r = ctx.finish(0)
        b.free()
    return r


full = gpu((seq, qual, off, lens))
want = po.accumulate_batch(seq, qual, off, lens, table)
bad = np.argwhere(full.rows != want.rows)
print("full batch: differing cells", len(bad), bad[:10].tolist())
# per-read first-hit according to the oracle
hits = []
for r in range(len(off)):
    one = po.accumulate_batch(*sub(r, r + 1), table)
    k = np.flatnonzero(one.rows[:, 96])
    hits.append(int(k[0]) if len(k) else -1)
hits = np.array(hits)
for p, c in bad.tolist():
    if c == 96:
        rs = np.flatnonzero(hits == p)
        print(f"pos {p}: got {full.rows[p, c]} want {want.rows[p, c]}; oracle reads with hit there: {rs[:40].tolist()}")
# which tile-size blocks fail on their own
RT = 93
for t0 in range(0, len(off), RT):
    g = gpu(sub(t0, min(t0 + RT, len(off))))
    w = po.accumulate_batch(*sub(t0, min(t0 + RT, len(off))), table)
    if not np.array_equal(g.rows, w.rows):
        d = np.argwhere(g.rows != w.rows)
        print("tile", t0 // RT, "reads", t0, "differs alone:", d.tolist()[:6])
        for r in range(t0, min(t0 + RT, len(off))):
            g1 = gpu(sub(r, r + 1))
            w1 = po.accumulate_batch(*sub(r, r + 1), table)
            if not np.array_equal(g1.rows, w1.rows):
                print("  read", r, "differs alone; off%16 =", int(off[r]) % 16, "seq", bytes(seq[off[r]:off[r] + 150]).decode())
                print("   got", np.flatnonzero(g1.rows[:, 96]).tolist(), "want", np.flatnonzero(w1.rows[:, 96]).tolist())
