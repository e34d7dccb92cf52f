"""N > 1 host logic on CPU: two gloo ranks each take their shard of the deterministic read stream, count it
(oracle -- there is no GPU here), sum the count arrays with a collective, and rank 0 must hold exactly the
single-process result.  This is the same partition + integer-sum structure the GPU path runs with one
ncclReduce (tests/test_gpu_multi.py covers that on hardware)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import capi, shard

N_READS = 6001


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ad, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, end = shard.shard_range(rank, world, N_READS)
    seq, qual, off, ln = capi.gen_reads(4, 1, begin, end - begin, 35, 300, 0.1)   # config-4 shaped shard
    res = po.accumulate_batch(seq, qual, off, ln, util.oracle_table() if ad else None)
    buf = torch.zeros(304 * 97 + 1, dtype=torch.int64)
    buf[: res.max_length * 97] = torch.from_numpy(res.rows.astype(np.int64).reshape(-1))
    buf[-1] = res.n_reads
    dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)      # the one exchange step of the path
    if rank == 0:
        np.save(out, buf.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("ad", [False, True], ids=["noad", "ad"])
def test_two_rank_shards_sum_to_the_whole(tmp_path, ad):
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, _free_port(), ad, out), nprocs=2, join=True)
    got = np.load(out)
    seq, qual, off, ln = capi.gen_reads(4, 1, 0, N_READS, 35, 300, 0.1)
    want = po.accumulate_batch(seq, qual, off, ln, util.oracle_table() if ad else None)
    rows = got[:-1].reshape(304, 97).astype(np.uint64)
    assert got[-1] == want.n_reads == N_READS
    ml = 1 + int(np.flatnonzero(rows[:, 95]).max())          # max_length derived from the length histogram
    assert ml == want.max_length and np.array_equal(rows[:ml], want.rows) and not rows[ml:].any()


def test_shard_range_partitions():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 1000003):
            spans = [shard.shard_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_rows_matches_whole():
    seq, qual, off, ln = capi.gen_reads(4, 1, 0, 3000, 35, 300, 0.1)
    whole = po.accumulate_batch(seq, qual, off, ln, None)
    parts = []
    for r in range(3):
        b, e = shard.shard_range(r, 3, 3000)
        s = capi.gen_reads(4, 1, b, e - b, 35, 300, 0.1)
        p = po.accumulate_batch(*s, None)
        parts.append((p.rows, p.max_length, p.n_reads))
    rows, ml, n = shard.merge_rows(parts, adapters_enabled=False)
    assert (ml, n) == (whole.max_length, whole.n_reads) and np.array_equal(rows, whole.rows)
