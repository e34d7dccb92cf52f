#!/usr/bin/env python
"""Development probe: bisects a random ragged batch down to the few reads on which the flat kernel (-a) and the
oracle disagree, and prints them.  usage: debug_flat.py [seed] [n] [lmin] [lmax]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import capi

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7000 + 35 * 400 + 300
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30011
lmin = int(sys.argv[3]) if len(sys.argv) > 3 else 35
lmax = int(sys.argv[4]) if len(sys.argv) > 4 else 300
table = util.oracle_table()
keys = table.keys()
seq, qual, off, lens = util.random_batch(seed, n, lmin, lmax, plant=0.3)


def sub(a, b):
    o0, o1 = int(off[a]), int(off[b - 1] + lens[b - 1])
    return seq[o0:o1], qual[o0:o1], (off[a:b] - off[a]).astype(np.uint32), lens[a:b]


def differs(a, b):
    batch = sub(a, b)
    with capi.Context(304, adapter_keys=keys, kernel=capi.KERNEL_FLAT) as ctx:
        ctx.accumulate_host(0, *batch)
        got = ctx.finish(0)
    want = po.accumulate_batch(*batch, table)
    return not (got.rows.shape == want.rows.shape and np.array_equal(got.rows, want.rows)), got, want


bad, got, want = differs(0, n)
print("whole batch differs:", bad)
if not bad:
    sys.exit(0)
# prefixes keep every offset, hence the chunking: the shortest prefix that differs ends with the read at fault
lo, hi = 0, n          # prefix of lo reads is fine, of hi reads differs
while hi - lo > 1:
    m = (lo + hi) // 2
    if differs(0, m)[0]:
        hi = m
    else:
        lo = m
r = hi - 1
_, got, want = differs(0, hi)
d = np.argwhere(got.rows != want.rows)
for p_, c in d[:10]:
    print("pos", p_, "col", c, "got", got.rows[p_, c], "want", want.rows[p_, c])
cb = (32 * 128 - 304 - 8) & ~63
print("culprit read", r, "offset", off[r], "len", lens[r], "offset mod 4", off[r] % 4, "chunk", off[r] // cb, "chunk bytes", cb,
      "offset in chunk window", off[r] % cb)
for k in range(max(0, r - 3), min(n, r + 3)):
    s_ = bytes(seq[off[k]: off[k] + lens[k]])
    one = po.accumulate_batch(*sub(k, k + 1), table)
    print(k, "off", off[k], "mod4", off[k] % 4, "chunk", off[k] // cb, "len", lens[k], "oracle kmer positions", np.flatnonzero(one.rows[:, 96]), s_[:60], s_[-20:])
