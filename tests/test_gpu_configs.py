"""BASELINE.json configs 3 and 4 as parity tests.

config 4: >= 5 M unpaired reads of 35..300 bp per GPU, with and without -a -- the whole batch against the oracle
          (the C restatement counts 5 M reads in a few seconds), plus the size-independent properties.
config 3: 200 M pairs = the config-2 files concatenated 20 x; counts are additive, so the expected arrays are exactly
          20 x the config-2 arrays.  Here: a config-2 batch per mate is counted 20 times (by the ranks of one node when
          there are several GPUs, tests/mp_config3_check.py) and compared with 20 x the oracle's counts of that batch.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _properties(r, n_reads, lmin, lmax):
    rows = r.rows
    assert r.n_reads == n_reads and lmin <= r.max_length <= lmax
    length = rows[:, capi.COL_LENGTH]
    assert int(length.sum()) == n_reads and not length[: lmin - 1].any()
    alive = n_reads - np.concatenate([[0], np.cumsum(length)[:-1]])   # reads at least p + 1 long
    assert np.array_equal(rows[:, capi.COL_CONTENT:capi.COL_CONTENT + 4].sum(axis=1), alive)
    assert np.array_equal(rows[:, :91].sum(axis=1), alive)
    assert int(rows[:, capi.COL_KMER].sum()) <= n_reads


@pytest.mark.parametrize("ad", [False, True], ids=["noad", "ad"])
def test_config4_ragged_5M_reads_vs_oracle(ad):
    n = 5_000_000
    table = util.oracle_table()
    keys = table.keys() if ad else None
    with capi.Context(304, n_mates=1, adapter_keys=keys, device_ids=[0]) as ctx:
        parts = []
        for first in range(0, n, 1_000_000):                 # 1 M-read device batches (170 MB each)
            parts.append(ctx.generate(4, 1, first, 1_000_000, 35, 300, 0.1))
        for b in parts:
            b.run(0)
        got = ctx.finish(0)
        launches = ctx.launch_count
        for b in parts:
            b.free()
    assert launches >= 5
    _properties(got, n, 35, 300)
    want = None
    for first in range(0, n, 1_000_000):
        w = po.accumulate_batch(*capi.gen_reads(4, 1, first, 1_000_000, 35, 300, 0.1), table if ad else None)
        if want is None:
            want = w
        else:
            want = capi.Result(want.rows + w.rows, max(want.max_length, w.max_length), want.n_reads + w.n_reads)
    util.assert_same(got, want, "config 4")


def test_config3_twenty_times_config2_single_gpu():
    """One GPU: the config-2 batch counted 20 times == 20 x (oracle counts of the batch), both mates."""
    pairs = 1_000_000
    table = util.oracle_table()
    with capi.Context(150, n_mates=2, adapter_keys=table.keys(), device_ids=[0]) as ctx:
        db = [ctx.generate(2, m + 1, 0, pairs, 150, 150, 0.1) for m in (0, 1)]
        for _ in range(20):
            db[0].run(0)
            db[1].run(1)
        got = [ctx.finish(0), ctx.finish(1)]
        for b in db:
            b.free()
    for m in (0, 1):
        w = po.accumulate_batch(*capi.gen_reads(2, m + 1, 0, pairs, 150, 150, 0.1), table)
        util.assert_same(got[m], capi.Result(w.rows * np.uint64(20), w.max_length, w.n_reads * 20), f"mate {m + 1}")


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_config3_sharded_over_the_node():
    n = min(_ngpu(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29641",
                        os.path.join(ROOT, "tests", "mp_config3_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and f"MP_CONFIG3_OK world={n}" in r.stdout, r.stdout[-3000:]
