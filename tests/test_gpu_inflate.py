"""On-device inflate of BGZF blocks: qb_bgzf_inflate() must give zlib's bytes for every kind
of DEFLATE block (stored, fixed codes, dynamic codes; long codes; matches that overlap their own output), check CRC-32
and ISIZE, and refuse damaged blocks; qb_bgzf_submit() (compressed bytes in, statistics out) must equal the oracle;
the `quack` program with QB_DEVICE_INFLATE=1 prints the same SVG and falls back to the host reader when refused.
BGZF layout: reference klib/bgzf.c:63-71 (header), 261-266 (trailer)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, capi, synth
from quack_b200.build import quack_bin

pytestmark = pytest.mark.gpu

QB_ERR_TEXT = -7
EOF_BLOCK = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


def bgzf(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, block=65280, eof=True, extra=b"") -> bytes:
    out = []
    for o in range(0, max(len(data), 1), block):
        part = data[o:o + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        raw = c.compress(part) + c.flush()
        xlen = 6 + len(extra)
        bsize = 12 + xlen + len(raw) + 8 - 1
        assert bsize < 65536
        out.append(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff]) + struct.pack("<H", xlen) + extra + b"BC" +
                   struct.pack("<HH", 2, bsize) + raw + struct.pack("<II", zlib.crc32(part), len(part)))
    return b"".join(out) + (EOF_BLOCK if eof else b"")


def _fastq_text(n, lmin, lmax, seed=1):
    seq, qual, off, ln = util.random_batch(seed, n, lmin, lmax, plant=0.2)
    out = []
    for r in range(n):
        o, l = int(off[r]), int(ln[r])
        out.append(b"@r%d\n" % r + seq[o:o + l].tobytes() + b"\n+\n" + qual[o:o + l].tobytes() + b"\n")
    return b"".join(out), (seq, qual, off, ln)


def _payloads():
    rng = np.random.default_rng(0)
    text, _ = _fastq_text(3000, 35, 151)
    return {
        "fastq": text,
        "random_bytes": rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes(),     # incompressible: stored blocks
        "runs": b"A" * 100_000 + b"AB" * 30_000 + b"ABC" * 10_000 + b"\n",                 # matches overlapping themselves
        "skewed": rng.choice(np.arange(256, dtype=np.uint8), size=150_000,
                             p=np.r_[0.9, np.full(255, 0.1 / 255)]).tobytes(),             # very long / very short codes
        "far_matches": (rng.integers(0, 256, size=30_000, dtype=np.uint8).tobytes()) * 4,   # distances up to 32 KiB
        "tiny": b"x",
        "empty": b"",
    }


@pytest.mark.parametrize("name", sorted(_payloads()))
def test_inflate_equals_zlib(name):
    data = _payloads()[name]
    with capi.Context(64, batch_bytes=1 << 20, ring_depth=2) as ctx:
        for level, strategy, block in ((0, zlib.Z_DEFAULT_STRATEGY, 65280), (1, zlib.Z_DEFAULT_STRATEGY, 65280),
                                       (6, zlib.Z_DEFAULT_STRATEGY, 65280), (9, zlib.Z_DEFAULT_STRATEGY, 30000),
                                       (6, zlib.Z_FIXED, 65280), (6, zlib.Z_HUFFMAN_ONLY, 50000), (6, zlib.Z_RLE, 777)):
            if level == 0 or name == "random_bytes":
                block = min(block, 60000)       # stored blocks: the payload does not shrink
            comp = bgzf(data, level, strategy, block)
            got = ctx.bgzf_inflate(comp, len(data) + 16)
            assert got == data, (name, level, strategy, block)


def test_inflate_generator_files_and_extra_fields(tmp_path):
    """The files the benchmark uses (tools/gen_fastq.cpp, zlib level 1), and blocks with more extra subfields."""
    p = str(tmp_path / "g.fq.gz")
    synth.write_fastq(p, 5, 1, 40_000, 150, 0.1, gz_level=1, bgzf=True)
    comp = open(p, "rb").read()
    text = b""
    rest = comp
    while rest:                                 # multi-member gzip: what gzread() delivers
        d = zlib.decompressobj(31)
        text += d.decompress(rest)
        rest = d.unused_data
    with capi.Context(64, batch_bytes=16 << 20, ring_depth=2) as ctx:
        assert ctx.bgzf_inflate(comp, len(text) + 16) == text
        data = _payloads()["fastq"]
        comp2 = bgzf(data, 6, extra=b"XY" + struct.pack("<H", 3) + b"abc")
        assert ctx.bgzf_inflate(comp2, len(data)) == data


def test_inflate_refuses_damaged_blocks():
    data = _payloads()["fastq"]
    comp = bytearray(bgzf(data, 6))
    with capi.Context(64, batch_bytes=1 << 20, ring_depth=2) as ctx:
        for at, what in ((100, "inside the deflate stream"), (len(comp) // 2, "middle"), (len(comp) - 40, "crc of the last block")):
            bad = bytearray(comp)
            bad[at] ^= 0x55
            with pytest.raises(capi.QbError) as e:
                ctx.bgzf_inflate(bytes(bad), len(data) + 16)
            assert e.value.code == QB_ERR_TEXT, what
        # ISIZE smaller / larger than what the stream holds
        first_total = struct.unpack_from("<H", comp, 16)[0] + 1
        for delta in (-1, 1):
            bad = bytearray(comp)
            isize = struct.unpack_from("<I", bad, first_total - 4)[0]
            struct.pack_into("<I", bad, first_total - 4, isize + delta)
            with pytest.raises(capi.QbError):
                ctx.bgzf_inflate(bytes(bad), len(data) + 16)
        with pytest.raises(capi.QbError):       # an ordinary gzip member is not a BGZF block
            ctx.bgzf_inflate(zlib.compress(data, 6, 31), len(data) + 16)
        assert ctx.bgzf_inflate(bytes(comp), len(data) + 16) == data


@pytest.mark.parametrize("shape", [(150, 150), (35, 300)], ids=lambda s: f"{s[0]}-{s[1]}")
@pytest.mark.parametrize("level", [1, 6])
def test_bgzf_submit_equals_the_oracle(shape, level):
    text, batch = _fastq_text(30_000, *shape, seed=shape[0])
    comp = bgzf(text, level)
    table = util.oracle_table()
    want = po.accumulate_batch(*batch, table)
    for max_chunk in (0, 300_000):               # whole slot at once / many chunks of a few blocks
        with capi.Context(320, adapter_keys=table.keys(), batch_bytes=1 << 20, ring_depth=3) as ctx:
            ctx.bgzf_accumulate(0, comp, max_chunk)
            assert ctx.text_status(0) == (30_000, 0)
            util.assert_same(ctx.finish(0), want, f"bgzf {shape} level {level} chunk {max_chunk}")


def test_bgzf_submit_damaged_is_refused():
    text, _ = _fastq_text(5000, 100, 100)
    comp = bytearray(bgzf(text, 6))
    comp[len(comp) // 2] ^= 0x10
    with capi.Context(128, batch_bytes=1 << 20, ring_depth=2) as ctx:
        ctx.bgzf_accumulate(0, bytes(comp))
        with pytest.raises(capi.QbError) as e:
            ctx.text_status(0)
        assert e.value.code == QB_ERR_TEXT


def _cli(args, env):
    e = dict(os.environ)
    e.update(env)
    return subprocess.run([quack_bin(), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)


def test_cli_device_inflate_same_svg(tmp_path):
    b1, b2, g2 = str(tmp_path / "b_1.fq.gz"), str(tmp_path / "b_2.fq.gz"), str(tmp_path / "g_2.fq.gz")
    synth.write_fastq(b1, 3, 1, 150_000, 150, 0.1, gz_level=1, bgzf=True)
    synth.write_fastq(b2, 3, 2, 150_000, 150, 0.1, gz_level=1, bgzf=True)
    synth.write_fastq(g2, 3, 2, 150_000, 150, 0.1, gz_level=1)
    args = ["-1", b1, "-2", b2, "-a", util.ADAPTER_FA, "-n", "inflate"]
    host = _cli(args, {"QB_BATCH_MB": "4"})
    dev = _cli(args, {"QB_BATCH_MB": "4", "QB_DEVICE_INFLATE": "1", "QB_VERBOSE": "1"})
    assert host.returncode == 0 and dev.returncode == 0, (host.stderr, dev.stderr)
    assert b"declined" not in dev.stderr and dev.stdout == host.stdout
    # unset: the device path is chosen by size (QB_DEVICE_INFLATE_MIN_MB, 1024 by default): not for these small files ...
    auto = _cli(args, {"QB_BATCH_MB": "4", "QB_VERBOSE": "1"})
    assert auto.returncode == 0 and b"on the device" not in auto.stderr and auto.stdout == host.stdout
    # ... but with the threshold lowered, and never with QB_DEVICE_INFLATE=0
    auto = _cli(args, {"QB_BATCH_MB": "4", "QB_VERBOSE": "1", "QB_DEVICE_INFLATE_MIN_MB": "1"})
    assert auto.returncode == 0 and b"on the device" in auto.stderr and auto.stdout == host.stdout
    off = _cli(args, {"QB_BATCH_MB": "4", "QB_VERBOSE": "1", "QB_DEVICE_INFLATE_MIN_MB": "1", "QB_DEVICE_INFLATE": "0"})
    assert off.returncode == 0 and b"on the device" not in off.stderr and off.stdout == host.stdout
    # one mate is ordinary gzip: the device frames it (text path), nothing is inflated on the device
    mixed = _cli(["-1", b1, "-2", g2, "-a", util.ADAPTER_FA, "-n", "inflate"], {"QB_DEVICE_INFLATE": "1", "QB_VERBOSE": "1"})
    assert mixed.returncode == 0 and mixed.stdout == host.stdout
    # a damaged block: the device refuses, the host reader reports what the reference would
    bad = bytearray(open(b1, "rb").read())
    bad[len(bad) // 2] ^= 0x04
    pbad = str(tmp_path / "bad.fq.gz")
    open(pbad, "wb").write(bytes(bad))
    a = _cli(["-u", pbad], {})
    b = _cli(["-u", pbad], {"QB_DEVICE_INFLATE": "1", "QB_VERBOSE": "1"})
    assert b"declined" in b.stderr and a.returncode == b.returncode and a.stdout == b.stdout
