"""Device-side record framing: qb_text_submit() takes decompressed FASTQ TEXT cut anywhere and
the device does what kseq_read() (klib/kseq.h:177-218) does for canonical 4-line records.  Checked against the oracle
on the same reads, with chunk cuts in every part of a record; text the device does not frame (multi-line records,
blank lines, CR LF, FASTA, garbage) must be REFUSED (QB_ERR_TEXT), never counted differently from kseq; the `quack`
program with QB_DEVICE_FRAMING=1 prints the same SVG as without, falling back to the host reader when refused."""
import os
import subprocess

import numpy as np
import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, capi, synth
from quack_b200.build import quack_bin

pytestmark = pytest.mark.gpu

QB_ERR_TEXT = -7


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


def _text_of(seq, qual, off, ln, names=None):
    out = []
    for r in range(len(off)):
        o, l = int(off[r]), int(ln[r])
        name = names[r] if names else b"r%d some comment" % r
        out.append(b"@" + name + b"\n" + seq[o:o + l].tobytes() + b"\n+\n" + qual[o:o + l].tobytes() + b"\n")
    return b"".join(out)


def _run_text(text, cuts, adapters=True, len_cap=320, batch_bytes=1 << 20, ring=3, kernel=capi.KERNEL_AUTO):
    table = util.oracle_table()
    with capi.Context(len_cap, adapter_keys=table.keys() if adapters else None, batch_bytes=batch_bytes,
                      ring_depth=ring, kernel=kernel) as ctx:
        ctx.text_accumulate(0, text, cuts)
        n, tail = ctx.text_status(0)
        res = ctx.finish(0)
        return res, n, tail, ctx.launch_count


@pytest.mark.parametrize("shape", [(150, 150), (35, 300), (1, 40), (100, 100)], ids=lambda s: f"{s[0]}-{s[1]}")
@pytest.mark.parametrize("adapters", [False, True], ids=["noad", "ad"])
def test_text_chunks_cut_anywhere_equal_the_oracle(shape, adapters):
    lmin, lmax = shape
    batch = util.random_batch(11 + lmin, 20_000, lmin, lmax, plant=0.2)
    text = _text_of(*batch)
    rng = np.random.default_rng(lmax)
    cuts = sorted(rng.integers(1, len(text), size=23).tolist())
    res, n, tail, launches = _run_text(text, cuts, adapters, batch_bytes=1 << 20)
    want = po.accumulate_batch(*batch, util.oracle_table() if adapters else None)
    assert (n, tail) == (20_000, 0)
    util.assert_same(res, want, f"text path {shape} adapters={adapters}")
    assert launches >= 1


def test_text_every_cut_position_of_a_small_file():
    """One cut at EVERY byte of a few records (inside the header, right behind a newline, inside the quality line...)."""
    batch = util.random_batch(5, 6, 20, 60, plant=0.5)
    names = [b"a", b"b/1 x", b"@@", b"+", b"c" * 70, b"d"]  # '@' and '+' inside header lines are legal
    text = _text_of(*batch, names=names)
    want = po.accumulate_batch(*batch, util.oracle_table())
    table = util.oracle_table()
    with capi.Context(64, adapter_keys=table.keys(), batch_bytes=1 << 16, ring_depth=3) as ctx:
        for cut in range(1, len(text)):
            ctx.reset(0)
            ctx.text_accumulate(0, text, [cut])
            assert ctx.text_status(0) == (6, 0), cut
            util.assert_same(ctx.finish(0), want, f"cut at {cut}")


def test_text_many_small_chunks_and_quality_lines_that_start_with_at_or_plus():
    rng = np.random.default_rng(3)
    reads = []
    for r in range(3000):
        l = int(rng.integers(1, 90))
        s = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=l).tobytes()
        q = bytearray((rng.integers(0, 42, size=l) + 33).astype(np.uint8).tobytes())
        if r % 3 == 0:
            q[0] = ord("@")           # quality 31: kseq reads the quality line by length, not by its first byte
        if r % 3 == 1:
            q[0] = ord("+")
        reads.append((s, bytes(q)))
    batch = util.pack(reads)
    text = _text_of(*batch)
    cuts = list(range(97, len(text), 1013))   # ~500 chunks, most of them smaller than a handful of records
    res, n, tail, _ = _run_text(text, cuts, True, len_cap=128, batch_bytes=1 << 18)
    assert (n, tail) == (3000, 0)
    util.assert_same(res, po.accumulate_batch(*batch, util.oracle_table()), "small chunks")


def test_text_truncated_tail_is_reported_not_counted():
    """The stream ends inside a record (no line end behind the last bytes): the complete records are counted, the
    left-over bytes are reported -- the `quack` program then lets the host reader decide (kseq returns -2 there)."""
    batch = util.random_batch(8, 500, 50, 50)
    text = _text_of(*batch)
    res, n, tail, _ = _run_text(text[:-20], [len(text) // 3], False, len_cap=64)
    seq, qual, off, ln = batch
    want = po.accumulate_batch(seq[:499 * 50], qual[:499 * 50], off[:499], ln[:499], None)
    assert n == 499 and tail > 0
    util.assert_same(res, want, "records in front of the truncated one")


BAD = {
    "multi_line_sequence": b"@a\nACGT\nACGT\n+\nIIIIIIII\n",
    "blank_line_between_records": b"@a\nACGT\n+\nIIII\n\n@b\nACGT\n+\nIIII\n",
    "crlf": b"@a\r\nACGT\r\n+\r\nIIII\r\n",
    "fasta": b">a\nACGT\n>b\nACGT\n>c\nAC\n>d\nAC\n",
    "garbage_in_front": b"hello\n@a\nACGT\n+\nIIII\n@b\nAC\n+\nII\n",
    "quality_shorter": b"@a\nACGT\n+\nIII\n@b\nACGT\n+\nIIII\n",
    "quality_longer": b"@a\nACGT\n+\nIIIII\n",
    "empty_sequence": b"@a\n\n+\n\n",
    "no_plus": b"@a\nACGT\n-\nIIII\n",
}


@pytest.mark.parametrize("name", sorted(BAD))
def test_text_that_is_not_canonical_is_refused(name):
    good = _text_of(*util.random_batch(1, 50, 30, 30))
    for text in (BAD[name], good + BAD[name], good + BAD[name] + good):
        with capi.Context(64, batch_bytes=1 << 16, ring_depth=2) as ctx:
            ctx.text_accumulate(0, text, [len(text) // 2])
            with pytest.raises(capi.QbError) as e:
                ctx.text_status(0)
            assert e.value.code == QB_ERR_TEXT, name


def test_text_two_mates_two_threads():
    import threading
    table = util.oracle_table()
    b1 = util.random_batch(21, 30_000, 150, 150, plant=0.2)
    b2 = util.random_batch(22, 30_000, 35, 151, plant=0.2)
    t1, t2 = _text_of(*b1), _text_of(*b2)
    errors = []

    def feed(ctx, mate, text):
        try:
            ctx.text_accumulate(mate, text, list(range(700_001, len(text), 700_001)))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    with capi.Context(160, n_mates=2, adapter_keys=table.keys(), batch_bytes=1 << 20, ring_depth=3) as ctx:
        th = [threading.Thread(target=feed, args=(ctx, 0, t1)), threading.Thread(target=feed, args=(ctx, 1, t2))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errors, errors
        assert ctx.text_status(0) == (30_000, 0) and ctx.text_status(1) == (30_000, 0)
        util.assert_same(ctx.finish(0), po.accumulate_batch(*b1, table), "mate 1")
        util.assert_same(ctx.finish(1), po.accumulate_batch(*b2, table), "mate 2")


def _cli(args, env):
    e = dict(os.environ)
    e.update(env)
    return subprocess.run([quack_bin(), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)


def test_cli_device_framing_same_svg(tmp_path, golden_dir):
    f1, f2 = str(tmp_path / "t_1.fq.gz"), str(tmp_path / "t_2.fq.gz")
    synth.write_fastq(f1, 3, 1, 120_000, 150, 0.1, gz_level=1)
    synth.write_fastq(f2, 3, 2, 120_000, 150, 0.1, gz_level=1, bgzf=True)
    args = ["-1", f1, "-2", f2, "-a", util.ADAPTER_FA, "-n", "framing"]
    host = _cli(args, {"QB_BATCH_MB": "4"})
    js = str(tmp_path / "s.json")
    dev = _cli(args, {"QB_BATCH_MB": "4", "QB_DEVICE_FRAMING": "1", "QB_VERBOSE": "1", "QB_STATS_JSON": js})
    assert host.returncode == 0 and dev.returncode == 0, (host.stderr, dev.stderr)
    assert b"declined" not in dev.stderr
    assert dev.stdout == host.stdout
    import json
    st = json.load(open(js))
    assert st["reads"] == 240_000 and st["bases"] == 240_000 * 150
    # the golden files with multi-line records / odd shapes: refused by the device, taken by the host reader
    for name in ("kat_t.fq", "kat_k.fq", "kat_p.dat", "rand_small.fq.gz"):
        p = os.path.join(golden_dir, name)
        if not os.path.exists(p):
            continue
        a = _cli(["-u", p, "-a", util.ADAPTER_FA], {})
        b = _cli(["-u", p, "-a", util.ADAPTER_FA], {"QB_DEVICE_FRAMING": "1", "QB_VERBOSE": "1"})
        assert a.returncode == b.returncode and a.stdout == b.stdout, name


TRICKY = {   # the streams of tests/test_host_cpu.py::test_reader_tricky_streams
    "only_garbage": b"no header here\n\n",
    "header_at_eof": b"@",
    "no_final_newline": b"@a\nACGT\n+\nIIII",
    "crlf": b"@a x\r\nACGT\r\n+\r\nIIII\r\n@b\r\nAC\r\n+\r\nII\r\n",
    "multi_line": b"@a\nAC\nGT\nAC\n+a\nII\nII\nII\n@b\nA\n+\nI\n",
    "qual_starts_with_at": b"@a\nACGT\n+\n@III\n@b\nAC\n+\n@@\n",
    "truncated_qual": b"@a\nACGTACGT\n+\nIII\n",
    "qual_too_long": b"@a\nACGT\n+\nIIIIII\n@b\nAC\n+\nII\n",
    "fasta_then_fastq": b">f\nACGT\n@a\nAC\n+\nII\n",
    "lone_cr_line": b"@a\n\r\nAC\n+\nIII\n",
    "plus_line_with_name": b"@a\nACGTACGTACGTA\n+a again\nIIIIIIIIIIIII\n@b\nACGTACGTACGTAC\n+\nIIIIIIIIIIIIII\n",
    "trailing_blank_lines": b"@a\nACGTACGTACGT\n+\nIIIIIIIIIIII\n\n\n",
}


@pytest.mark.parametrize("name", sorted(TRICKY))
def test_cli_device_framing_on_tricky_streams_equals_host_reader(name, tmp_path):
    """Whatever the device makes of a stream -- frames it, or refuses and lets the host reader take it -- the program's
    output is the one the kseq-exact host reader gives (which tests/test_host_cpu.py pins to the reference parse)."""
    good = _text_of(*util.random_batch(4, 40, 30, 60))
    alone = name in ("no_final_newline", "qual_starts_with_at", "multi_line", "only_garbage")  # (every run pays CUDA start-up)
    for i, data in enumerate(((TRICKY[name],) if alone else ()) + (good + TRICKY[name],)):
        p = tmp_path / f"{name}_{i}.fq"
        p.write_bytes(data)
        a = _cli(["-u", str(p), "-a", util.ADAPTER_FA], {})
        b = _cli(["-u", str(p), "-a", util.ADAPTER_FA], {"QB_DEVICE_FRAMING": "1"})
        assert (a.returncode, a.stdout) == (b.returncode, b.stdout), (name, i, a.stderr, b.stderr)
