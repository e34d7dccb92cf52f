"""Host side of the drop-in (C, CPU only): the batching record reader against the reference reader's
golden parses, adapter keys from a FASTA file, and the transform+draw restatement against the SVGs the
unmodified reference binary printed (tests/golden/svg) -- byte for byte."""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, capi
from quack_b200.build import quack_bin


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = {"kat_t": "kat_t.fq", "kat_k": "kat_k.fq", "framing": "framing.fq", "rand_small": "rand_small.fq.gz",
         "kat_p": "kat_p.dat"}   # kat_p = the reference's own parser fixture, klib/test/kseq_test.dat


def test_reader_matches_reference_parse(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "golden_parse.json")))
    for name in ("framing", "kat_k", "kat_p"):
        recs, rc = capi.parse_records(os.path.join(golden_dir, FILES[name]))
        assert rc == gold[name]["rc"]
        assert [[s.decode("latin1"), q.decode("latin1") if q else None] for s, q in recs] == gold[name]["records"]


@pytest.mark.parametrize("name", sorted(FILES))
def test_reader_equals_oracle_reader(name, golden_dir):
    path = os.path.join(golden_dir, FILES[name])
    assert capi.parse_records(path) == po.parse_records(path)


def test_reader_tricky_streams(tmp_path):
    cases = {
        "empty": b"",
        "only_garbage": b"no header here\n\n",
        "header_at_eof": b"@",
        "no_final_newline": b"@a\nACGT\n+\nIIII",
        "crlf": b"@a x\r\nACGT\r\n+\r\nIIII\r\n@b\r\nAC\r\n+\r\nII\r\n",
        "multi_line": b"@a\nAC\nGT\nAC\n+a\nII\nII\nII\n@b\nA\n+\nI\n",
        "qual_starts_with_at": b"@a\nACGT\n+\n@III\n@b\nAC\n+\n@@\n",
        "truncated_qual": b"@a\nACGTACGT\n+\nIII\n",
        "qual_too_long": b"@a\nACGT\n+\nIIIIII\n@b\nAC\n+\nII\n",
        "fasta_then_fastq": b">f\nACGT\n@a\nAC\n+\nII\n",
        "zero_length": b"@a\n\n+\n\n@b\nAC\n+\nII\n",
        "lone_cr_line": b"@a\n\r\nAC\n+\nIII\n",
    }
    for name, data in cases.items():
        p = tmp_path / (name + ".fq")
        p.write_bytes(data)
        assert capi.parse_records(str(p)) == po.parse_records(str(p)), name
    g = tmp_path / "multi_member.fq.gz"
    g.write_bytes(gzip.compress(b"@a\nACGT\n+\nIIII\n") + gzip.compress(b"@b\nAC\n+\nII\n"))
    recs, rc = capi.parse_records(str(g))
    assert recs == [(b"ACGT", b"IIII"), (b"AC", b"II")] and rc == -1


def test_batches_cover_the_file_in_order(golden_dir):
    path = os.path.join(golden_dir, "rand_small.fq.gz")
    recs, _ = po.parse_records(path)
    batches, status = capi.read_batches(path, cap_bytes=7000, cap_reads=40)
    assert status == -1 and len(batches) > 20
    got = []
    for seq, qual, off, ln, ml in batches:
        assert len(off) <= 40 and len(seq) <= 7000 and (len(ln) == 0 or ml == ln.max())
        assert np.array_equal(off[1:], np.cumsum(ln)[:-1].astype(np.uint32)) and (len(off) == 0 or off[0] == 0)
        got += [(seq[o: o + l].tobytes(), qual[o: o + l].tobytes()) for o, l in zip(off, ln)]
    assert got == recs
    # truncated stream: everything before the bad record is delivered, then the stream is over (quack.c:193)
    batches, status = capi.read_batches(os.path.join(golden_dir, "framing.fq"), 1 << 16, 100)
    assert status == -2 and sum(len(b[2]) for b in batches) == 3


def test_read_adapters_from_fasta(golden_dir, adapters_fa, tmp_path):
    gold = np.load(os.path.join(golden_dir, "golden_adapter_keys.npy"))
    keys = capi.read_adapters(adapters_fa)
    assert len(keys) == 769 and np.array_equal(np.unique(keys), gold)
    gz = tmp_path / "a.fa.gz"
    gz.write_bytes(gzip.compress(open(adapters_fa, "rb").read()))
    assert np.array_equal(capi.read_adapters(str(gz)), keys)


def _gold_result(raw, name, tag):
    ml, n = raw[f"{name}.{tag}.meta"]
    return capi.Result(raw[f"{name}.{tag}.rows"], int(ml), int(n))


def _golden_svg(golden_dir, name):
    p = os.path.join(golden_dir, "svg", name + ".svg")
    return open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()


RENDER = {  # golden svg -> (first fixture, second fixture, adapters, name)
    "u_kat_t": ("kat_t", None, False, None),
    "u_kat_t_ad": ("kat_t", None, True, None),
    "u_kat_t_ad_name": ("kat_t", None, True, "tiny"),
    "u_kat_k_ad": ("kat_k", None, True, None),
    "pe_kat_t_kat_k_ad_name": ("kat_t", "kat_k", True, "pair"),
    "u_rand": ("rand_small", None, False, None),
    "u_rand_ad": ("rand_small", None, True, None),
    "pe_rand_kat_k": ("rand_small", "kat_k", False, None),
    "pe_rand_rand_ad_name": ("rand_small", "rand_small", True, "rand x2"),
}


@pytest.mark.parametrize("svg", sorted(RENDER))
def test_render_is_byte_identical_to_reference(svg, golden_dir, tmp_path):
    first, second, ad, name = RENDER[svg]
    raw = np.load(os.path.join(golden_dir, "golden_raw.npz"))
    tag = "ad" if ad else "noad"
    out = str(tmp_path / "o.svg")
    capi.render_svg(_gold_result(raw, first, tag), _gold_result(raw, second, tag) if second else None, ad, name, out)
    got = open(out, "rb").read()
    want = _golden_svg(golden_dir, svg)
    md5 = json.load(open(os.path.join(golden_dir, "svg", "index.json")))["md5"][svg]
    assert hashlib.md5(want).hexdigest() == md5
    if got != want:
        gl, wl = got.split(b"\n"), want.split(b"\n")
        for i, (a, b) in enumerate(zip(gl, wl)):
            assert a == b, f"line {i + 1}: {a[:160]!r} != {b[:160]!r}"
        assert len(gl) == len(wl)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built")
def test_render_vs_live_reference_random_and_binning(tmp_path):
    """Fresh inputs through the reference binary itself, incl. reads > 3000 bp (100-bp binning) and a
    phred64-looking file (encoding detection)."""
    rng = np.random.default_rng(12)
    specs = {"long": (40, 2500, 3400, 2, 41), "phred64": (300, 50, 120, 31, 71), "mid": (500, 1, 480, 0, 60)}  # <= 500: the reference keeps averages[500] on the stack
    for name, (n, lmin, lmax, qlo, qhi) in specs.items():
        p = str(tmp_path / f"{name}.fq")
        with open(p, "wb") as f:
            for i in range(n):
                l = int(rng.integers(lmin, lmax + 1))
                s = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=l).tobytes()
                q = (rng.integers(qlo, qhi + 1, size=l).astype(np.uint8) + 33).tobytes()
                f.write(b"@r%d\n%s\n+\n%s\n" % (i, s, q))
        for ad in (None, util.ADAPTER_FA):
            want = po.ref_svg(["-u", p] + (["-a", ad] if ad else []))
            res = po.ref_read_fastq(p, ad)   # raw arrays from the reference; only the renderer is under test
            out = str(tmp_path / "o.svg")
            capi.render_svg(capi.Result(res.rows, res.max_length, res.n_reads), None, ad is not None, None, out)
            assert open(out, "rb").read() == want, (name, ad)


def test_cli_usage_and_exit_codes():
    q = quack_bin()
    run = lambda *a: subprocess.run([q, *a], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    r = run("-V")
    assert (r.returncode, r.stdout) == (0, b"quack 1.1.1\n")
    r = run()
    assert r.returncode == 1 and r.stdout.startswith(b"Usage: quack [OPTION...]\nquack -- A FASTQ") and \
        r.stdout.endswith(b"Try `quack --help' or `quack --usage' for more information.\n")
    r = run("-x", "y", "-u")  # even argc: nothing parsed
    assert r.returncode == 1
    r = run("--bogus", "1")
    assert r.returncode == 1 and r.stderr.startswith(b"Usage: quack") and r.stdout == b""
    if po.have_ref():
        for args in ([], ["-V"], ["--help"], ["-?"], ["--usage"], ["--bogus", "1"], ["-u", "x", "-1", "a", "-2", "b"],
                     ["-n", "name"], ["-1", "only_forward"]):
            ref = subprocess.run([po.REF_BIN, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            got = run(*args)
            assert (got.returncode, got.stdout, got.stderr) == (ref.returncode, ref.stdout, ref.stderr), args


# ---- parallel BGZF decode: same bytes to the framing code as gzread() ----

def _fastq_blob(n=3000, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        l = int(rng.integers(1, 400))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=l))
        q = bytes((rng.integers(2, 42, size=l) + 33).astype(np.uint8))
        out.append(b"@r%d some comment\n%s\n+\n%s\n" % (i, s, q))
    return b"".join(out)


@pytest.mark.parametrize("block", [64, 1000, 65280])
@pytest.mark.parametrize("threads", [1, 2, 8])
def test_bgzf_pool_equals_gzread_and_oracle(block, threads, tmp_path):
    """Records straddle block boundaries (tiny blocks), every thread count gives the serial result."""
    from quack_b200 import synth
    blob = _fastq_blob()
    b, g = tmp_path / "r.fq.bgz", tmp_path / "r.fq.gz"
    b.write_bytes(synth.bgzf_bytes(blob, block=block))
    g.write_bytes(gzip.compress(blob, 1))
    assert gzip.decompress(b.read_bytes()) == blob            # a BGZF file is a multi-member gzip file
    want = capi.parse_records(str(g), 1)
    assert want[1] == -1 and len(want[0]) == 3000
    assert capi.parse_records(str(b), threads) == want
    assert po.parse_records(str(b)) == want
    # the batching entry point, small batches
    bb, st = capi.read_batches(str(b), 20000, 100, threads)
    gb, gst = capi.read_batches(str(g), 20000, 100, 1)
    assert st == gst == -1 and len(bb) == len(gb)
    for x, y in zip(bb, gb):
        assert all(np.array_equal(u, v) for u, v in zip(x[:4], y[:4])) and x[4] == y[4]


def test_bgzf_pool_edge_cases(tmp_path):
    from quack_b200 import synth
    blob = _fastq_blob(400, seed=6)
    good = synth.bgzf_bytes(blob, block=3000)
    cases = {}
    cases["empty"] = (synth.bgzf_bytes(b""), 0, -1)
    cases["no_eof_marker"] = (good[: -len(synth.BGZF_EOF)], 400, -1)
    p = tmp_path / "x.bgz"
    for name, (data, n, rc) in cases.items():
        p.write_bytes(data)
        for t in (1, 4):
            recs, st = capi.parse_records(str(p), t)
            assert (len(recs), st) == (n, rc), (name, t)
    # a block cut short, a flipped payload byte: stream error (-3, where kseq_read() returns -3 too), the records
    # of the intact blocks in front of it are kept
    first = int.from_bytes(good[16:18], "little") + 1
    second = int.from_bytes(good[first + 16: first + 18], "little") + 1
    t = tmp_path / "plain.fq"
    t.write_bytes(blob)
    want, _ = capi.parse_records(str(t), 1)
    broken = {
        "truncated": good[: first + second - 5],
        "flipped": good[: first + 30] + bytes([good[first + 30] ^ 0x55]) + good[first + 31:],
    }
    for name, data in broken.items():
        p.write_bytes(data)
        recs, st = capi.parse_records(str(p), 4)
        assert st == -3, name
        if recs and recs[-1][1] is None:   # kseq semantics: a record cut inside its sequence comes back as FASTA
            recs = recs[:-1]
        assert 0 < len(recs) < len(want) and recs == want[: len(recs)], name   # a prefix of the true record list


def test_bgzf_followed_by_other_members_reads_like_gzread(tmp_path):
    """A BGZF file that continues with an ordinary gzip member, or ends in bytes that are no gzip member at all:
    gzread() (the reference, quack.c:187) decodes the member and ignores the garbage; so does the pool."""
    from quack_b200 import synth
    blob = _fastq_blob(400, seed=7)
    good = synth.bgzf_bytes(blob, block=3000)
    first = int.from_bytes(good[16:18], "little") + 1
    tail = b"@z\nACGTACGTAC\n+\nIIIIIIIIII\n"
    cases = {
        "plain_member_mid": good[:first] + gzip.compress(tail),
        "plain_member_after_all": good[: -len(synth.BGZF_EOF)] + gzip.compress(tail) + gzip.compress(tail),
        "bgzf_after_plain_member": good[:first] + gzip.compress(tail) + synth.bgzf_bytes(tail),
        "trailing_garbage": good + b"this is not gzip",
        "trailing_zeros": good + bytes(64),
    }
    p = tmp_path / "m.bgz"
    for name, data in cases.items():
        p.write_bytes(data)
        want = capi.parse_records(str(p), 1)        # gzread path
        assert len(want[0]) > 0 and want[1] in (-1, -2), name   # (a member boundary may cut a record: -2, as gzread)
        assert po.parse_records(str(p)) == want, name
        for t in (2, 8):
            assert capi.parse_records(str(p), t) == want, (name, t)
            bb, st = capi.read_batches(str(p), 20000, 100, t)
            n_fastq = len([r for r in want[0] if r[1] is not None])
            assert st in (-1, -2, -5) and sum(len(b[2]) for b in bb) in (n_fastq, len(want[0])), (name, t)


@pytest.mark.parametrize("kind", ["plain", "gz", "bgzf"])
def test_reader_on_a_pipe(kind, tmp_path):
    """Non-seekable input (FIFO, process substitution, /dev/stdin): opened exactly once, nothing is lost to the
    BGZF probe, whatever the thread count (the reference reads pipes through one gzopen, quack.c:187)."""
    import threading
    from quack_b200 import synth
    blob = _fastq_blob(1000, seed=9)
    data = {"plain": blob, "gz": gzip.compress(blob, 1), "bgzf": synth.bgzf_bytes(blob)}[kind]
    f = tmp_path / "reg"
    f.write_bytes(data)
    want = capi.parse_records(str(f), 1)
    assert len(want[0]) == 1000 and want[1] == -1
    for threads in (0, 4):
        fifo = str(tmp_path / f"fifo{threads}")
        os.mkfifo(fifo)

        def feed():
            with open(fifo, "wb") as w:
                w.write(data)
        th = threading.Thread(target=feed)
        th.start()
        got = capi.parse_records(fifo, threads)
        th.join()
        assert got == want, threads
    # the program itself on /dev/stdin never gets as far as CUDA without a GPU; the reader is what matters here
    r = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path.insert(0, %r); from quack_b200 import capi; "
                        "r, st = capi.parse_records('/dev/stdin', 4); print(len(r), st)" % ROOT],
                       input=data, stdout=subprocess.PIPE, check=True)
    assert r.stdout.split() == [b"1000", b"-1"]


def test_pools_are_chosen_by_container(tmp_path):
    """Plain text keeps the gzread() path whatever the thread count; ordinary gzip takes the member pool, BGZF the
    block pool; one thread: gzread() for everything, like the reference."""
    p = tmp_path / "a.fq"
    p.write_bytes(b"@a\nACGT\n+\nIIII\n")
    assert capi.decode_throughput(str(p), 8)["threads"] == 1
    g = tmp_path / "a.fq.gz"
    g.write_bytes(gzip.compress(b"@a\nACGT\n+\nIIII\n"))
    d = capi.decode_throughput(str(g), 8)
    assert d["threads"] == 8 and d["reads"] == 1
    assert capi.decode_throughput(str(g), 1)["threads"] == 1
    from quack_b200 import synth
    b = tmp_path / "a.fq.bgz"
    b.write_bytes(synth.bgzf_bytes(b"@a\nACGT\n+\nIIII\n"))
    d = capi.decode_throughput(str(b), 8)
    assert d["threads"] == 8 and d["reads"] == 1 and d["bases"] == 4


def _gzip_with_name(data: bytes) -> bytes:
    """A member with FNAME and FCOMMENT header fields (RFC 1952), as `gzip file` writes them."""
    import struct
    import zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    return (b"\x1f\x8b\x08\x18" + struct.pack("<I", 1_700_000_000) + b"\x00\x03" + b"reads.fq\x00" + b"a comment\x00" + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF))


@pytest.mark.parametrize("threads", [2, 8])
def test_gzip_member_pool_equals_gzread(threads, tmp_path):
    """Ordinary (multi-member) gzip through the speculative member pool: the same records, the same final status
    as gzread() -- the reference's decoder (quack.c:187) -- on every shape of file."""
    blob = _fastq_blob(3000, seed=11)
    cut = [0, 1000, 1001, 50_000, 50_017, 300_000, len(blob)]
    members = [gzip.compress(blob[a:b], 1) for a, b in zip(cut, cut[1:])]
    fake = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + b"ACGT" * 12
    stored = gzip.compress(b"@s\nACGT" + fake + b"\n+\nIIII" + b"I" * len(fake) + b"\n", 0)   # the fake header verbatim in the file
    tail_rec = b"@t\nACGTACGT\n+\nIIIIIIII\n"
    cases = {
        "one_member": gzip.compress(blob, 6),
        "many_members": b"".join(members),
        "empty_members_between": members[0] + gzip.compress(b"") + members[1] + gzip.compress(b"") + b"".join(members[2:]),
        "trailing_garbage": b"".join(members) + b"garbage that starts no member",
        "trailing_zeros": b"".join(members) + bytes(100),
        "header_lookalike_inside_a_member": members[0] + stored + gzip.compress(tail_rec),
        "truncated_last_member": b"".join(members)[:-40],
        "truncated_mid": b"".join(members[:4]) + members[4][: len(members[4]) // 2],
        "with_name_and_comment_fields": members[0] + _gzip_with_name(blob[1000:50_000]),
        "member_with_exotic_os_byte": members[0] + members[1][:9] + b"\x63" + members[1][10:] + b"".join(members[2:]),
    }
    p = tmp_path / "m.fq.gz"
    for name, data in cases.items():
        p.write_bytes(data)
        want = capi.parse_records(str(p), 1)        # gzread path
        assert po.parse_records(str(p)) == want, name
        got = capi.parse_records(str(p), threads)
        assert got == want, (name, len(got[0]), got[1], len(want[0]), want[1])
        bb, st = capi.read_batches(str(p), 50_000, 400, threads)
        gb, gst = capi.read_batches(str(p), 50_000, 400, 1)
        assert st == gst and len(bb) == len(gb), name
        for x, y in zip(bb, gb):
            assert all(np.array_equal(u, v) for u, v in zip(x[:4], y[:4])), name


def test_damaged_gzip_streams(tmp_path):
    """A flipped byte inside a member: every decoder stops with -3 (kseq_read()'s stream error).  zlib notices the
    damage some way behind the flipped byte and the bytes in between inflate to something else than was written, so
    the records handed out before are NOT a prefix of the true list -- but they are the same for every decoder as far
    as it gets.  How far depends on the request size against zlib: the gzread() path asks for kseq's 16 KiB at a
    time and so stops exactly where the reference's reader stops (checked against it when oracle/_ref is there); the
    pools hand out every byte that inflates in front of the point where zlib gives up, never fewer records."""
    blob = _fastq_blob(3000, seed=11)
    cut = [0, 1000, 1001, 50_000, 50_017, 300_000, len(blob)]
    members = [gzip.compress(blob[a:b], 1) for a, b in zip(cut, cut[1:])]
    data = b"".join(members[:4]) + members[4][:2000] + bytes([members[4][2000] ^ 0x5A]) + members[4][2001:] + members[5]
    p = tmp_path / "flip.fq.gz"
    p.write_bytes(data)

    def complete(recs):
        return recs[:-1] if recs and recs[-1][1] is None else recs   # (a record cut inside its sequence comes back as FASTA)
    serial, st = capi.parse_records(str(p), 1)
    assert st == -3 and len(complete(serial)) > 100
    if po.have_ref():
        ref, rst = po.ref_parse_records(str(p), 1 << 26)
        assert rst == -3 and ref == serial
    for threads in (2, 8):
        pool, st = capi.parse_records(str(p), threads)
        assert st == -3 and len(complete(pool)) >= len(complete(serial))
        assert complete(pool)[: len(complete(serial))] == complete(serial)


def test_gzip_member_pool_big_members(tmp_path):
    """Members of several MiB (the benchmark generator writes 64 MiB of text per member): chunks cross the pool in
    order, the speculative workers run ahead of the consumer."""
    blob = _fastq_blob(40_000, seed=12)      # ~16 MB of text
    third = len(blob) // 3
    data = b"".join(gzip.compress(blob[i: i + third + 7], 1) for i in range(0, len(blob), third + 7))
    p = tmp_path / "big.fq.gz"
    p.write_bytes(data)
    want = capi.parse_records(str(p), 1)
    assert len(want[0]) == 40_000 and want[1] == -1
    for t in (2, 4, 16):
        assert capi.parse_records(str(p), t) == want, t


