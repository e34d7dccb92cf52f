"""Parity of the CUDA path (through the C-ABI) with the oracle: bit-exact count arrays.

Every test here needs a B200; run with `pytest -m gpu`.  The oracle is the checker only.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from quack_b200 import capi
import qb_testutil as util

pytestmark = pytest.mark.gpu

KERNELS = [capi.KERNEL_SIMPLE, capi.KERNEL_FUSED]
SMEM_KERNELS = [capi.KERNEL_FUSED]
KNAME = {capi.KERNEL_SIMPLE: "simple", capi.KERNEL_FUSED: "fused"}


@pytest.fixture(scope="module")
def table():
    return util.oracle_table()


@pytest.fixture(scope="module")
def keys(table):
    return table.keys()


# ---- period kernel (v5): uniform-length batches only ----
PERIOD_LENS = [32, 33, 36, 50, 51, 64, 76, 100, 101, 126, 128, 130, 150, 151, 200, 248, 250, 256]


@pytest.mark.parametrize("l", PERIOD_LENS)
@pytest.mark.parametrize("ad", [False, True], ids=["noad", "ad"])
@pytest.mark.parametrize("resident", [False, True], ids=["stream", "resident"])
def test_period_uniform(l, ad, resident, table, keys):
    """Every supported period geometry (3..5 steps, 1 and 2 histogram blocks), a read count that leaves a
    remainder for the other kernels, planted adapters, N bases."""
    batch = util.random_batch(1000 + l, 20011, l, l, plant=0.3)
    got = run_gpu(batch, 256, keys if ad else None, capi.KERNEL_PERIOD, resident=resident)
    util.assert_same(got, po.accumulate_batch(*batch, table if ad else None), f"period {l}")


@pytest.mark.parametrize("l", [100, 150, 151])
def test_period_all_byte_values(l, table, keys):
    rng = np.random.default_rng(l)
    n = 6000
    seq = rng.integers(0, 256, size=n * l).astype(np.uint8)
    qual = rng.integers(0, 256, size=n * l).astype(np.uint8)
    off = (np.arange(n) * l).astype(np.uint32)
    lens = np.full(n, l, dtype=np.uint32)
    want = po.accumulate_batch(seq, qual, off, lens, table)
    got = run_gpu((seq, qual, off, lens), 256, keys, capi.KERNEL_PERIOD)
    util.assert_same(got, want, "period all bytes")
    assert got.invalid == want.n_invalid_qual > 0


@pytest.mark.parametrize("qlo,qhi", [(0, 90), (31, 71), (44, 49)], ids=["full", "phred64", "edge"])
def test_period_score_ranges(qlo, qhi, table, keys):
    batch = util.random_batch(78, 9000, 150, 150, qlo=qlo, qhi=qhi, plant=0.2)
    got = run_gpu(batch, 150, keys, capi.KERNEL_PERIOD)
    util.assert_same(got, po.accumulate_batch(*batch, table), f"period scores {qlo}-{qhi}")


def test_period_adapter_every_position(table, keys):
    """First-hit semantics with the hit planted at every position of a 150-bp read, a second adapter behind
    it, hits ending on the last base, poly-A tails (many anchor hits per read: the queue overflows)."""
    ad = b"AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"
    reads = []
    l = 150
    for rep in range(3):
        for at in range(0, l):
            s = bytearray((b"C" if rep != 1 else b"A") * l)
            m = min(len(ad), l - at)
            s[at: at + m] = ad[:m]
            if at + 51 <= l:
                s[at + 40: at + 40 + 11] = ad[:11]
            reads.append((bytes(s), b"I" * l))
    batch = util.pack(reads * 5)
    got = run_gpu(batch, 150, keys, capi.KERNEL_PERIOD)
    util.assert_same(got, po.accumulate_batch(*batch, table), "period adapter positions")


def test_period_u16_flush(monkeypatch, table, keys):
    """One CTA, > 65535 reads: the u16 counters are flushed in time."""
    monkeypatch.setenv("QB_FUSED_GRID", "1")
    l = 150
    n = 150000
    rng = np.random.default_rng(9)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * l)]
    qual = np.full(n * l, 33 + 37, dtype=np.uint8)   # one score bin: every read hits the same counters
    off = (np.arange(n) * l).astype(np.uint32)
    lens = np.full(n, l, dtype=np.uint32)
    got = run_gpu((seq, qual, off, lens), 150, keys, capi.KERNEL_PERIOD, resident=True)
    util.assert_same(got, po.accumulate_batch(seq, qual, off, lens, table), "period flush")


def test_period_rejects_ragged(keys):
    batch = util.random_batch(3, 100, 100, 150)
    with pytest.raises(capi.QbError):
        run_gpu(batch, 150, keys, capi.KERNEL_PERIOD)


# ---- flat kernel (v6): ragged batches of back-to-back reads of 16..320 bp ----
FLAT_SHAPES = [(35, 300, 304), (16, 40, 40), (16, 304, 304), (100, 151, 151), (150, 150, 150), (17, 17, 64), (299, 304, 304),
               (64, 64, 64), (48, 96, 96)]


@pytest.mark.parametrize("lmin,lmax,cap", FLAT_SHAPES)
@pytest.mark.parametrize("ad", [False, True], ids=["noad", "ad"])
@pytest.mark.parametrize("resident", [False, True], ids=["stream", "resident"])
def test_flat_ragged(lmin, lmax, cap, ad, resident, table, keys):
    """Every read-length mix the flat kernel takes: reads that start at every byte phase of a 16-byte unit and of a
    32-bit word (straddling words), chunk boundaries inside units, planted adapters, N bases."""
    batch = util.random_batch(7000 + lmin * 400 + lmax, 30011, lmin, lmax, plant=0.3)
    got = run_gpu(batch, cap, keys if ad else None, capi.KERNEL_FLAT, resident=resident)
    util.assert_same(got, po.accumulate_batch(*batch, table if ad else None), f"flat {lmin}-{lmax}")


@pytest.mark.parametrize("chunk", ["512", "1000", "2048"])
def test_flat_small_chunks(chunk, monkeypatch, table, keys):
    """Small byte windows: many chunk boundaries, chunks of one or two reads, reads longer than half a window."""
    monkeypatch.setenv("QB_FLAT_CHUNK", chunk)
    for lmin, lmax in ((35, 300), (16, 60), (280, 304)):
        batch = util.random_batch(int(chunk) + lmin, 20000, lmin, lmax, plant=0.3)
        got = run_gpu(batch, 304, keys, capi.KERNEL_FLAT, resident=True)
        util.assert_same(got, po.accumulate_batch(*batch, table), f"flat chunk {chunk} {lmin}-{lmax}")


def test_flat_longest_reads(table):
    """Up to 320 bp without -a; with -a the packed codes and first hits take the room of the last 16 positions."""
    batch = util.random_batch(320, 20000, 16, 320)
    util.assert_same(run_gpu(batch, 320, None, capi.KERNEL_FLAT), po.accumulate_batch(*batch, None), "flat 16-320")


def test_flat_tiny_batches(table, keys):
    for n in (1, 2, 3, 31, 33, 100):
        batch = util.random_batch(n, n, 16, 304, plant=0.5)
        got = run_gpu(batch, 304, keys, capi.KERNEL_FLAT)
        util.assert_same(got, po.accumulate_batch(*batch, table), f"flat n={n}")


def test_flat_all_byte_values_and_score_ranges(table, keys):
    rng = np.random.default_rng(15)
    lens = rng.integers(16, 200, size=5000).astype(np.uint32)
    off = np.zeros(len(lens), dtype=np.uint32)
    off[1:] = np.cumsum(lens[:-1], dtype=np.uint64).astype(np.uint32)
    total = int(lens.sum())
    seq = rng.integers(0, 256, size=total).astype(np.uint8)
    qual = rng.integers(0, 256, size=total).astype(np.uint8)
    want = po.accumulate_batch(seq, qual, off, lens, table)
    got = run_gpu((seq, qual, off, lens), 200, keys, capi.KERNEL_FLAT)
    util.assert_same(got, want, "flat all bytes")
    assert got.invalid == want.n_invalid_qual > 0
    for qlo, qhi in ((0, 90), (31, 71), (44, 49), (60, 62)):
        batch = util.random_batch(78, 8000, 40, 151, qlo=qlo, qhi=qhi, plant=0.2)
        util.assert_same(run_gpu(batch, 151, keys, capi.KERNEL_FLAT), po.accumulate_batch(*batch, table), f"flat scores {qlo}-{qhi}")


def test_flat_adapter_first_hit_every_position(table, keys):
    """The first-hit rule with the adapter planted at every position of reads of three lengths, a second adapter
    further down (only the first may count), hits that end on the last base, windows that would span two reads."""
    ad = b"AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"
    reads = []
    for l in (30, 61, 150):
        for at in range(0, l):
            s = bytearray(b"C" * l)
            m = min(len(ad), l - at)
            s[at: at + m] = ad[:m]
            if at + 51 <= l:
                s[at + 40: at + 40 + 11] = ad[:11]
            reads.append((bytes(s), b"I" * l))
    reads.append((b"CCCCCCCCCCCCGATCGGAAGA", b"I" * 22))       # hit ends on the last base
    reads.append((b"GATCGGAAGACCCCCC", b"I" * 16))             # a window that starts in the read in front must not count
    reads.append((b"CCCCCCCCCCCGATCG", b"I" * 16))
    reads.append((b"GAAGACCCCCCCCCCC", b"I" * 16))
    batch = util.pack(reads)
    util.assert_same(run_gpu(batch, 150, keys, capi.KERNEL_FLAT), po.accumulate_batch(*batch, table), "flat first hit")
    dimers = [(ad[:30] * 5, b"I" * 150)] * 400                 # every anchor hits: the hit queue overflows and is drained in rounds
    batch = util.pack(dimers)
    util.assert_same(run_gpu(batch, 150, keys, capi.KERNEL_FLAT), po.accumulate_batch(*batch, table), "flat dimers")


def test_flat_u16_flush(monkeypatch, table, keys):
    monkeypatch.setenv("QB_FUSED_GRID", "2")
    batch = util.random_batch(32, 300000, 30, 50, plant=0.05)
    got = run_gpu(batch, 64, keys, capi.KERNEL_FLAT, resident=True)
    util.assert_same(got, po.accumulate_batch(*batch, table), "flat flush")


def test_flat_rejects_other_batches(keys):
    with pytest.raises(capi.QbError):     # reads shorter than a 16-byte unit
        run_gpu(util.random_batch(3, 100, 5, 40), 64, keys, capi.KERNEL_FLAT)
    seq, qual, off, lens = util.random_batch(4, 100, 50, 60)
    off = off.copy()
    off[50:] += 7                          # a gap between two reads
    pad = np.zeros(7, dtype=np.uint8)
    seq = np.concatenate([seq[: off[50] - 7], pad, seq[off[50] - 7:]])
    qual = np.concatenate([qual[: off[50] - 7], pad + 40, qual[off[50] - 7:]])
    with pytest.raises(capi.QbError):
        run_gpu((seq, qual, off, lens), 64, keys, capi.KERNEL_FLAT)
    with capi.Context(64, adapter_keys=keys) as ctx:   # AUTO still counts it (fused kernel), bit-exactly
        ctx.accumulate_host(0, seq, qual, off, lens)
        got = ctx.finish(0)
        assert ctx.flat_launch_count == 0
    util.assert_same(got, po.accumulate_batch(seq, qual, off, lens, util.oracle_table()), "gap batch, auto")


def test_auto_takes_the_flat_kernel_for_ragged_batches(table, keys):
    batch = util.random_batch(5, 50000, 35, 300, plant=0.1)
    with capi.Context(304, adapter_keys=keys) as ctx:
        ctx.accumulate_host(0, *batch)
        got = ctx.finish(0)
        assert ctx.flat_launch_count >= 1 and ctx.period_launch_count == 0
    util.assert_same(got, po.accumulate_batch(*batch, table), "auto ragged")


def run_gpu(batch, len_cap, keys, kernel, resident=False, **kw):
    seq, qual, off, lens = batch
    with capi.Context(len_cap, adapter_keys=keys, kernel=kernel, **kw) as ctx:
        if resident:
            b = ctx.upload(seq, qual, off, lens, max_len=int(lens.max()) if len(lens) else 0)
            b.run(0)
            res = ctx.finish(0)
            b.free()
        else:
            ctx.accumulate_host(0, seq, qual, off, lens)
            res = ctx.finish(0)
        res.invalid = ctx.invalid_quality_count(0)
    return res


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
@pytest.mark.parametrize("name", ["kat_t", "kat_k", "framing", "rand_small"])
@pytest.mark.parametrize("ad", [False, True], ids=["noad", "ad"])
def test_golden_fixtures(name, ad, kernel, golden_dir, keys):
    """Reference-generated raw arrays (tests/golden/golden_raw.npz) reproduced by the kernels."""
    files = {"kat_t": "kat_t.fq", "kat_k": "kat_k.fq", "framing": "framing.fq", "rand_small": "rand_small.fq.gz"}
    recs, _ = po.parse_records(os.path.join(golden_dir, files[name]))
    batch = util.pack(recs)
    got = run_gpu(batch, 304, keys if ad else None, kernel)
    gold = np.load(os.path.join(golden_dir, "golden_raw.npz"))
    ml, n = gold[f"{name}.{'ad' if ad else 'noad'}.meta"]
    assert (got.max_length, got.n_reads) == (int(ml), int(n))
    assert np.array_equal(got.rows, gold[f"{name}.{'ad' if ad else 'noad'}.rows"])


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
@pytest.mark.parametrize("ad", [False, True], ids=["noad", "ad"])
@pytest.mark.parametrize("resident", [False, True], ids=["stream", "resident"])
def test_fixed_150(kernel, ad, resident, table, keys):
    batch = util.random_batch(11, 20000, 150, 150, plant=0.2)
    got = run_gpu(batch, 150, keys if ad else None, kernel, resident=resident)
    util.assert_same(got, po.accumulate_batch(*batch, table if ad else None), "fixed150")


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
@pytest.mark.parametrize("lmin,lmax,cap", [(1, 40, 40), (35, 300, 304), (0, 25, 64), (9, 12, 150), (300, 300, 300)])
def test_ragged(kernel, lmin, lmax, cap, table, keys):
    batch = util.random_batch(lmin * 1000 + lmax, 30000, lmin, lmax, plant=0.3)
    got = run_gpu(batch, cap, keys, kernel)
    util.assert_same(got, po.accumulate_batch(*batch, table), f"ragged {lmin}-{lmax}")


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
def test_all_byte_values(kernel, table, keys):
    """Every byte value in both streams: both sides define the bytes the reference leaves undefined
    the same way, and count out-of-range quality bytes instead of corrupting neighbours."""
    rng = np.random.default_rng(5)
    n, l = 4000, 97
    seq = rng.integers(0, 256, size=n * l).astype(np.uint8)
    qual = rng.integers(0, 256, size=n * l).astype(np.uint8)
    off = (np.arange(n) * l).astype(np.uint32)
    lens = np.full(n, l, dtype=np.uint32)
    want = po.accumulate_batch(seq, qual, off, lens, table)
    got = run_gpu((seq, qual, off, lens), 128, keys, kernel)
    util.assert_same(got, want, "all bytes")
    assert got.invalid == want.n_invalid_qual > 0


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
@pytest.mark.parametrize("qlo,qhi", [(0, 90), (31, 71), (44, 49), (60, 62)], ids=["full", "phred64", "edge", "high"])
def test_score_ranges(kernel, qlo, qhi, table, keys):
    """Scores beyond the shared-memory window (Phred > 46, e.g. phred64 files) take the exact slow path."""
    batch = util.random_batch(77, 8000, 100, 151, qlo=qlo, qhi=qhi)
    got = run_gpu(batch, 151, keys, kernel)
    util.assert_same(got, po.accumulate_batch(*batch, table), f"scores {qlo}-{qhi}")


def test_adapter_first_hit_every_position(table, keys):
    """First-hit semantics with the hit planted at every position, two adapters per read
    (only the first may count), hits ending on the last base (no count) and windows spanning reads."""
    ad = b"AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"
    reads = []
    for l in (30, 61, 150):
        for at in range(0, l):
            s = bytearray(b"C" * l)
            m = min(len(ad), l - at)
            s[at: at + m] = ad[:m]
            if at + 51 <= l:
                s[at + 40: at + 40 + 11] = ad[:11]
            reads.append((bytes(s), b"I" * l))
    reads.append((b"CCCCCCCCCCCCGATCGGAAGA", b"I" * 22))       # hit ends on the last base
    reads.append((b"GATCG", b"IIIII"))                        # window would span into the next read
    reads.append((b"GAAGACCCCCCCCCCC", b"I" * 16))
    batch = util.pack(reads)
    want = po.accumulate_batch(*batch, table)
    for kernel in KERNELS:
        util.assert_same(run_gpu(batch, 150, keys, kernel), want, KNAME[kernel])


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
def test_two_mates_many_small_batches(kernel, table, keys):
    """Ring of tiny slots: many submits, slot reuse, two independent accumulators."""
    b1 = util.random_batch(21, 12000, 20, 150, plant=0.1)
    b2 = util.random_batch(22, 9000, 150, 150, plant=0.1)
    with capi.Context(150, n_mates=2, adapter_keys=keys, kernel=kernel, batch_bytes=64 << 10, batch_reads=700,
                      ring_depth=2) as ctx:
        ctx.accumulate_host(0, *b1)
        ctx.accumulate_host(1, *b2)
        ctx.accumulate_host(0, *b2)
        r0, r1 = ctx.finish(0), ctx.finish(1)
        assert ctx.launch_count > 30
    w1, w2 = po.accumulate_batch(*b1, table), po.accumulate_batch(*b2, table)
    util.assert_same(r1, w2, "mate 1")
    assert r0.n_reads == w1.n_reads + w2.n_reads and r0.max_length == 150
    rows = w2.rows.copy()
    rows[: w1.max_length] += w1.rows
    assert np.array_equal(r0.rows, rows)


@pytest.mark.parametrize("kernel", SMEM_KERNELS, ids=KNAME.get)
def test_u16_flush_path(kernel, table, keys, monkeypatch):
    """More than 65535 reads through one CTA forces the mid-launch flush of the packed u16 counters."""
    monkeypatch.setenv("QB_FUSED_GRID", "2")
    batch = util.random_batch(31, 300000, 30, 50, plant=0.05)
    got = run_gpu(batch, 64, keys, kernel, resident=True)
    util.assert_same(got, po.accumulate_batch(*batch, table), "flush")


def test_linearity_and_generator_full_size(keys, table):
    """Config-2 shape at a size the oracle cannot chew quickly: 2 M generated 150-bp reads.  Checked by
    properties (running the batch twice doubles every count; per-position content and score sums equal
    the number of reads that long) and against the oracle on the first 100 k reads."""
    n = 2_000_000
    with capi.Context(150, adapter_keys=keys, kernel=capi.KERNEL_FUSED) as ctx:
        b = ctx.generate(2, 1, 0, n, 150, 150, 0.1)
        b.run(0)
        r1 = ctx.finish(0)
        b.run(0)
        r2 = ctx.finish(0)
        b.free()
        assert r1.n_reads == n and r2.n_reads == 2 * n and r1.max_length == 150
        assert np.array_equal(r2.rows, 2 * r1.rows)
        assert np.all(r1.rows[:, 91:95].sum(axis=1) == n) and np.all(r1.rows[:, :91].sum(axis=1) == n)
        assert r1.rows[149, 95] == n
        # ~10 % planted read-through + ~3.4 % chance hits on random sequence
        assert 0.10 * n < r1.rows[:, 96].sum() < 0.16 * n
    seq, qual, off, lens = capi.gen_reads(2, 1, 0, 100_000, 150, 150, 0.1)
    want = po.accumulate_batch(seq, qual, off, lens, table)
    for kernel in KERNELS:
        util.assert_same(run_gpu((seq, qual, off, lens), 150, keys, kernel, resident=True), want, "gen 100k")


def test_capacity_errors(keys):
    batch = util.random_batch(3, 100, 60, 60)
    with capi.Context(40, adapter_keys=keys) as ctx:
        with pytest.raises(capi.QbError) as e:
            ctx.accumulate_host(0, *batch)
        assert e.value.code == -4
    with capi.Context(150, adapter_keys=keys) as ctx:     # resident batch bypasses the host check: kernel flags it
        b = ctx.upload(*batch, max_len=60)
        ctx2 = capi.Context(40, adapter_keys=keys)
        b2 = ctx2.upload(*batch, max_len=40)
        b2.run(0)
        with pytest.raises(capi.QbError):
            ctx2.finish(0)
        b2.free()
        ctx2.close()
        b.run(0)
        assert ctx.finish(0).n_reads == 100
        b.free()


def test_empty_and_reset(keys):
    with capi.Context(150, adapter_keys=keys) as ctx:
        r = ctx.finish(0)
        assert (r.max_length, r.n_reads, r.rows.shape) == (0, 0, (0, 97))
        batch = util.random_batch(4, 500, 150, 150)
        ctx.accumulate_host(0, *batch)
        assert ctx.finish(0).n_reads == 500
        ctx.reset(0)
        assert ctx.finish(0).n_reads == 0


@pytest.mark.parametrize("kernel", KERNELS, ids=KNAME.get)
def test_large_adapter_set(kernel):
    """An adapter file with thousands of k-mers does not fit the shared-memory key table: the kernels then
    confirm filter hits against the 2^20-bit map in L2.  Dense set -> many first hits."""
    rng = np.random.default_rng(99)
    recs = [rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=60).tobytes() for _ in range(120)]
    table = po.AdapterTable.from_records(recs)
    keys = table.keys()
    assert len(keys) > 3000
    batch = util.random_batch(123, 20000, 40, 150, plant=0.3, plant_seq=recs[5])
    got = run_gpu(batch, 150, keys, kernel)
    util.assert_same(got, po.accumulate_batch(*batch, table), "large set")


def test_auto_kernel_is_planned_per_batch(table, keys):
    """len_cap only sizes the accumulator: a context opened for 65536-bp reads (the CLI default) still
    runs short-read batches on the fused kernel, and falls back to the simple kernel only for a batch
    whose longest read is beyond the shared-memory histogram."""
    short = util.random_batch(41, 20000, 35, 300, plant=0.2)
    long_ = util.random_batch(42, 300, 200, 900, plant=0.2)
    with capi.Context(65536, adapter_keys=keys) as ctx:
        ctx.accumulate_host(0, *short)
        res = ctx.finish(0)
        assert ctx.kernel_counts == (0, ctx.launch_count) and ctx.launch_count >= 1
        util.assert_same(res, po.accumulate_batch(*short, table), "short reads, big len_cap")
        n_fused = ctx.launch_count
        ctx.accumulate_host(0, *long_)
        res = ctx.finish(0)
        assert ctx.kernel_counts == (ctx.launch_count - n_fused, n_fused) and ctx.launch_count > n_fused
    want = po.accumulate_batch(*long_, table)
    w1 = po.accumulate_batch(*short, table)
    rows = want.rows.copy()
    rows[: w1.max_length] += w1.rows
    assert res.n_reads == want.n_reads + w1.n_reads and res.max_length == want.max_length
    assert np.array_equal(res.rows, rows)


@pytest.mark.parametrize("kernel", SMEM_KERNELS, ids=KNAME.get)
def test_max_len_promise_is_enforced(kernel, keys):
    """A batch submitted with a max_len smaller than its longest read is rejected loudly, not mis-counted."""
    batch = util.random_batch(43, 200, 100, 100)
    with capi.Context(304, adapter_keys=keys, kernel=kernel) as ctx:
        b = ctx.upload(*batch, max_len=50)
        b.run(0)
        with pytest.raises(capi.QbError):
            ctx.finish(0)
        b.free()


def test_adapter_dimers_overflow_the_hit_queue(table, keys):
    """Reads made of adapter sequence end to end: every anchor passes the filter, the per-tile queue of
    anchor hits overflows and the kernel falls back to testing every window of the tile exactly."""
    ads = [r for r in util.adapter_records() if len(r) >= 30]
    rng = np.random.default_rng(7)
    reads = []
    for i in range(6000):
        s = b"".join(ads[int(k)] for k in rng.integers(0, len(ads), size=6))
        l = 150 if i % 3 else int(rng.integers(40, 151))
        reads.append((s[:l], bytes(rng.integers(35, 74, size=l).astype(np.uint8))))
    batch = util.pack(reads)
    want = po.accumulate_batch(*batch, table)
    assert want.rows[:, 96].sum() > 5000
    for kernel in KERNELS:
        util.assert_same(run_gpu(batch, 150, keys, kernel), want, KNAME[kernel])
    uniform = util.pack([(s[:150].ljust(150, b"A"), q[:150].ljust(150, b"I")) for s, q in reads])
    for kernel in SMEM_KERNELS:
        util.assert_same(run_gpu(uniform, 150, keys, kernel, resident=True),
                         po.accumulate_batch(*uniform, table), "uniform dimers " + KNAME[kernel])
