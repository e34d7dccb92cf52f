"""The pinned ring under several host threads: a slot handed out by qb_acquire() belongs to the caller until its
qb_submit(), whatever the other threads do meanwhile.  The `quack` program runs one reader
thread per mate on ONE ring (reference: the two read_fastq() calls of quack.c:911-921), so mates that decode at
different speeds -- BGZF next to gzip, a file next to a slow pipe -- are exactly this case.  Also here: the
accumulators that grow with the longest read seen (reads longer than the old 65536-bp default)."""
import ctypes as C
import gzip
import os
import subprocess
import threading
import time

import numpy as np
import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, capi, synth
from quack_b200.build import quack_bin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


def _fill_and_submit(ctx, mate, batches, delay, errors):
    """One 'reader thread': acquire, (sleep), fill, submit -- the order quack_main.c:mate_thread uses."""
    try:
        for seq, qual, off, ln in batches:
            b = ctx.acquire()
            if delay:
                time.sleep(delay)            # a slow decoder: the slot stays HELD meanwhile
            n, nb = len(off), len(seq)
            C.memmove(b.seq, seq.ctypes.data, nb)
            C.memmove(b.qual, qual.ctypes.data, nb)
            C.memmove(b.offset, off.ctypes.data, 4 * n)
            C.memmove(b.length, ln.ctypes.data, 4 * n)
            if delay:
                time.sleep(delay / 4)
            ctx.submit(b, mate, n, nb, int(ln.max()))
    except Exception as e:  # noqa: BLE001
        errors.append(e)


@pytest.mark.parametrize("ring", [2, 3])
def test_two_threads_different_speed_share_one_ring(ring):
    """Thread A cycles through many batches while thread B sits on its slot: with a ring that only tracks
    `submitted` (round 1), A's k-th acquire is handed B's slot and both write the same pinned buffers."""
    table = util.oracle_table()
    fast = [util.random_batch(100 + i, 3000, 150, 150, plant=0.2) for i in range(40)]
    slow = [util.random_batch(900 + i, 2500, 35, 140, plant=0.2) for i in range(5)]
    errors = []
    with capi.Context(150, n_mates=2, adapter_keys=table.keys(), batch_bytes=1 << 20, batch_reads=8192,
                      ring_depth=ring) as ctx:
        ta = threading.Thread(target=_fill_and_submit, args=(ctx, 0, fast, 0.0, errors))
        tb = threading.Thread(target=_fill_and_submit, args=(ctx, 1, slow, 0.08, errors))
        tb.start()
        time.sleep(0.02)                     # B holds its first slot before A starts cycling
        ta.start()
        ta.join()
        tb.join()
        assert not errors, errors
        got = [ctx.finish(0), ctx.finish(1)]
    for g, batches, what in ((got[0], fast, "fast mate"), (got[1], slow, "slow mate")):
        want = po.accumulate_batch(*batches[0], table)
        for b in batches[1:]:
            w = po.accumulate_batch(*b, table)
            rows = np.zeros((max(want.max_length, w.max_length), capi.ROW), dtype=np.uint64)
            rows[: want.max_length] += want.rows
            rows[: w.max_length] += w.rows
            want = capi.Result(rows, max(want.max_length, w.max_length), want.n_reads + w.n_reads)
        util.assert_same(g, want, what)


def test_submit_needs_an_acquired_slot():
    with capi.Context(64, batch_bytes=1 << 16, ring_depth=2) as ctx:
        b = ctx.acquire()
        ctx.submit(b, 0, 0, 0)
        with pytest.raises(capi.QbError):
            ctx.submit(b, 0, 0, 0)           # the slot went back to the ring with the first submit


def test_four_threads_ring_of_two():
    """More threads than slots: qb_acquire blocks until somebody submits, nobody shares a slot."""
    table = util.oracle_table()
    batches = [[util.random_batch(7 * t + i, 1500, 80, 150, plant=0.1) for i in range(6)] for t in range(4)]
    errors = []
    with capi.Context(150, n_mates=1, adapter_keys=table.keys(), batch_bytes=1 << 19, batch_reads=4096,
                      ring_depth=2) as ctx:
        ths = [threading.Thread(target=_fill_and_submit, args=(ctx, 0, batches[t], 0.01 * t, errors)) for t in range(4)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert not errors, errors
        got = ctx.finish(0)
    n = sum(len(b[2]) for bs in batches for b in bs)
    assert got.n_reads == n
    allb = [b for bs in batches for b in bs]
    seq = np.concatenate([b[0] for b in allb])
    qual = np.concatenate([b[1] for b in allb])
    ln = np.concatenate([b[3] for b in allb])
    off = np.zeros(len(ln), dtype=np.uint64)
    off[1:] = np.cumsum(ln[:-1], dtype=np.uint64)
    util.assert_same(got, po.accumulate_batch(seq, qual, off.astype(np.uint32), ln, table), "4 threads")


# ------------------------------------------------------------------------------------------ the program

def _run(args, env=None, **kw):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([quack_bin(), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e, **kw)


def _reference_svg(args, tmp_path, tag):
    """What the unmodified reference prints for these arguments (oracle/_ref when it travelled with the snapshot),
    else the oracle's arrays through the host renderer."""
    if po.have_ref():
        return po.ref_svg(args)
    opt = dict(zip(args[::2], args[1::2]))
    t = util.oracle_table() if "-a" in opt else None
    a, b = po.read_fastq(opt["-1"], t), po.read_fastq(opt["-2"], t)
    out = str(tmp_path / (tag + ".svg"))
    capi.render_svg(capi.Result(a.rows, a.max_length, a.n_reads), capi.Result(b.rows, b.max_length, b.n_reads),
                    t is not None, opt.get("-n"), out)
    return open(out, "rb").read()


@pytest.fixture(scope="module")
def mate_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("mates")
    f = {}
    for mate in (1, 2):
        for kind, kw in (("plain", {}), ("gz", {"gz_level": 1}), ("bgzf", {"gz_level": 1, "bgzf": True})):
            p = str(d / (f"m{mate}.{kind}.fq" + ("" if kind == "plain" else ".gz")))
            synth.write_fastq(p, 2, mate, 150_000, 150, 0.1, **kw)
            f[(mate, kind)] = p
    return f


@pytest.mark.parametrize("k1,k2", [("bgzf", "gz"), ("gz", "bgzf"), ("plain", "gz"), ("bgzf", "plain")])
@pytest.mark.parametrize("ring", ["2", "3"])
def test_cli_mates_of_different_decode_speed(k1, k2, ring, mate_files, tmp_path):
    """-1 and -2 decoded at very different speeds (BGZF pool with 8 threads ~ 2 GB/s of text, gzread ~ 0.25 GB/s,
    plain text faster still) through small slots: the SVG equals the reference's, byte for byte."""
    args = ["-1", mate_files[(1, k1)], "-2", mate_files[(2, k2)], "-a", util.ADAPTER_FA, "-n", "mixed"]
    r = _run(args, {"QB_RING": ring, "QB_BATCH_MB": "1", "QUACK_DECODE_THREADS": "8"})
    assert r.returncode == 0, r.stderr
    ref_args = ["-1", mate_files[(1, "gz")], "-2", mate_files[(2, "gz")], "-a", util.ADAPTER_FA, "-n", "mixed"]
    assert r.stdout == _reference_svg(ref_args, tmp_path, "mixed")


def test_cli_one_mate_through_a_slow_pipe(mate_files, tmp_path):
    """-2 arrives through a FIFO that a throttled writer feeds (a pipeline / process substitution): the fast mate
    laps the slow one many times on the shared ring; and the BGZF probe must not eat the head of a pipe."""
    fifo = str(tmp_path / "slow.fq.gz")
    os.mkfifo(fifo)
    data = open(mate_files[(2, "gz")], "rb").read()

    def feed():
        with open(fifo, "wb") as w:
            for i in range(0, len(data), 1 << 18):
                w.write(data[i: i + (1 << 18)])
                time.sleep(0.01)
    th = threading.Thread(target=feed)
    th.start()
    args = ["-1", mate_files[(1, "bgzf")], "-2", fifo, "-a", util.ADAPTER_FA]
    r = _run(args, {"QB_RING": "2", "QB_BATCH_MB": "1", "QUACK_DECODE_THREADS": "8"})
    th.join()
    assert r.returncode == 0, r.stderr
    ref_args = ["-1", mate_files[(1, "gz")], "-2", mate_files[(2, "gz")], "-a", util.ADAPTER_FA]
    assert r.stdout == _reference_svg(ref_args, tmp_path, "pipe")


def test_cli_stdin_and_process_substitution(mate_files, tmp_path):
    args = ["-u", "/dev/stdin", "-a", util.ADAPTER_FA]
    r = _run(args, input=open(mate_files[(1, "gz")], "rb").read())
    assert r.returncode == 0, r.stderr
    want = _run(["-u", mate_files[(1, "gz")], "-a", util.ADAPTER_FA])
    assert want.returncode == 0 and r.stdout == want.stdout
    sh = subprocess.run(["bash", "-c", f"{quack_bin()} -u <(cat {mate_files[(1, 'plain')]}) -a {util.ADAPTER_FA}"],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert sh.returncode == 0 and sh.stdout == want.stdout, sh.stderr


def test_cli_truncated_input_is_reported(mate_files, tmp_path):
    """A damaged stream ends the counting like it does in the reference (quack.c:193) -- with a warning on stderr."""
    data = open(mate_files[(1, "gz")], "rb").read()
    cut = tmp_path / "cut.fq.gz"
    cut.write_bytes(data[: len(data) // 2])
    r = _run(["-u", str(cut)])
    # (the cut usually falls inside a record: the reader then sees a record without its quality line first, which the
    # reference would count with stale quality bytes -- quack.c:203 -- so the two SVGs are not compared here)
    assert r.returncode == 0 and b"warning" in r.stderr and b"records in front of it" in r.stderr
    assert r.stdout.startswith(b"<svg")


# ------------------------------------------------------------------------------------------ long reads

def _long_read_file(path, lens, seed=5):
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        for i, l in enumerate(lens):
            s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=l))
            q = bytes((rng.integers(2, 42, size=l) + 33).astype(np.uint8))
            f.write(b"@long%d\n%s\n+\n%s\n" % (i, s, q))


def test_cli_reads_longer_than_65536_bp(tmp_path):
    """The accumulators grow with the longest read (512 rows at start): ultra-long reads run with the defaults,
    through the reference's 100-bp binning (quack.c:234-262).  Up to 50 000 bp the reference binary itself is the
    judge; beyond that its fixed averages[500] (quack.c:301) overflows, so the oracle's arrays through the host
    renderer are."""
    p = str(tmp_path / "ont.fq")
    _long_read_file(p, [200, 41_000, 3100, 45_000, 90, 12_345])
    r = _run(["-u", p])
    assert r.returncode == 0, r.stderr
    if po.have_ref():
        assert r.stdout == po.ref_svg(["-u", p])
    p2 = str(tmp_path / "ultra.fq")
    _long_read_file(p2, [200, 70_000, 3100, 150_000, 90, 65_537], seed=6)
    r = _run(["-u", p2])
    assert r.returncode == 0, r.stderr
    a = po.read_fastq(p2, None)
    out = str(tmp_path / "ultra.svg")
    capi.render_svg(capi.Result(a.rows, a.max_length, a.n_reads), None, False, None, out)
    assert r.stdout == open(out, "rb").read()
    r2 = _run(["-u", p2], {"QB_LEN_CAP": "100000"})
    assert r2.returncode == 2 and b"QB_LEN_CAP" in r2.stderr


def test_accumulators_grow_and_keep_counts():
    """Short batches, then a longer one, then short ones again: growth keeps everything counted so far."""
    table = util.oracle_table()
    parts = [util.random_batch(1, 5000, 100, 150, plant=0.1), util.random_batch(2, 300, 400, 1500),
             util.random_batch(3, 5000, 150, 150, plant=0.1), util.random_batch(4, 40, 3000, 9000),
             util.random_batch(5, 2000, 35, 300)]
    with capi.Context(1 << 20, n_mates=2, adapter_keys=table.keys(), batch_bytes=4 << 20) as ctx:
        for mate in (0, 1):
            for b in (parts if mate == 0 else parts[::-1]):
                ctx.accumulate_host(mate, *b)
        got = [ctx.finish(0), ctx.finish(1)]
    seq = np.concatenate([b[0] for b in parts])
    qual = np.concatenate([b[1] for b in parts])
    ln = np.concatenate([b[3] for b in parts])
    off = np.zeros(len(ln), dtype=np.uint64)
    off[1:] = np.cumsum(ln[:-1], dtype=np.uint64)
    want = po.accumulate_batch(seq, qual, off.astype(np.uint32), ln, table)
    util.assert_same(got[0], want, "mate 0")
    util.assert_same(got[1], want, "mate 1 (reverse order)")
