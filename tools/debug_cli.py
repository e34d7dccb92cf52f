#!/usr/bin/env python
"""Debug aid: config-2 shaped mates through the streaming path vs the oracle."""
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
for mate in (1, 2):
    seq, qual, off, lens = capi.gen_reads(2, mate, 0, 200_000, 150, 150, 0.1)
    want = po.accumulate_batch(seq, qual, off, lens, table)
    for bb in (8 << 20, 64 << 20):
        with capi.Context(65536, adapter_keys=keys, batch_bytes=bb) as ctx:
            ctx.accumulate_host(0, seq, qual, off, lens)
            got = ctx.finish(0)
            print("mate", mate, "batch", bb >> 20, "launches", ctx.launch_count, ctx.kernel_counts,
                  "equal", np.array_equal(got.rows, want.rows), "n", got.n_reads, want.n_reads)
            if not np.array_equal(got.rows, want.rows):
                bad = np.argwhere(got.rows != want.rows)
                print("  cells", len(bad), [(int(p), int(c), int(got.rows[p, c]), int(want.rows[p, c])) for p, c in bad[:12]])
