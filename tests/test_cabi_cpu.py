"""CPU-side checks of the boundary: the library loads, exports every symbol include/quack_b200.h
declares, refuses to run without a GPU (no fallback), and its host helpers agree with the oracle."""
import os
import re

import numpy as np
import pytest

from oracle import pyoracle as po
from quack_b200 import build, capi
import qb_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "quack_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    L = capi.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.QbError) as e:
        capi.Context(150)
    assert "no CPU fallback" in str(e.value)


def test_base_code_equals_oracle_on_all_bytes():
    assert [capi.base_code(c) for c in range(256)] == [po.base_code(c) for c in range(256)]


def test_adapter_keys_equal_reference_table(golden_dir):
    keys = np.concatenate([capi.adapter_record_keys(r) for r in util.adapter_records()])
    gold = np.load(os.path.join(golden_dir, "golden_adapter_keys.npy"))
    assert len(keys) == 769 and np.array_equal(np.unique(keys), gold)   # 769 windows, 333 keys
    assert len(capi.adapter_record_keys(b"ACGTACGTAC")) == 0


def test_generator_is_deterministic_and_shardable():
    a = capi.gen_reads(2, 1, 0, 3000, 150, 150, 0.1)
    b = capi.gen_reads(2, 1, 1000, 1000, 150, 150, 0.1)
    assert np.array_equal(a[0][1000 * 150: 2000 * 150], b[0]) and np.array_equal(a[1][1000 * 150: 2000 * 150], b[1])
    m2 = capi.gen_reads(2, 2, 0, 3000, 150, 150, 0.1)
    assert not np.array_equal(a[0], m2[0])
    q = a[1].astype(int) - 33
    assert q.min() >= 2 and q.max() <= 41
    assert set(np.unique(a[0])) <= set(b"ACGTN")
    v = capi.gen_reads(4, 1, 0, 5000, 35, 300, 0.0)
    assert v[3].min() >= 35 and v[3].max() <= 300 and v[2][-1] + v[3][-1] == len(v[0])
    # read-through adapters make the oracle's adapter panel non-trivial: ~10 % + chance hits
    r = po.accumulate_batch(*a, util.oracle_table())
    assert 0.09 * 3000 < r.rows[:, 96].sum() < 0.18 * 3000


@pytest.mark.parametrize("l", [32, 33, 36, 50, 51, 64, 76, 100, 101, 126, 128, 130, 150, 151, 200, 248, 250, 256])
@pytest.mark.parametrize("ad", [0, 1])
def test_period_layout_is_a_valid_counter_map(l, ad):
    """Host side of the period kernel (no GPU needed): the plan covers whole reads and 16-byte tiles, and
    the bank-conflict solver returns an injective position -> (block, column, half) map that is never
    worse than the natural layout (column = position / 4)."""
    import ctypes as C
    L = capi.lib()
    L.qb_period_layout.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    info, slot = (C.c_uint32 * 7)(), (C.c_uint8 * 256)()
    assert L.qb_period_layout(l, ad, info, slot) == 0
    k, wp, steps, ppt, rpt, stages, warps = list(info)
    assert k * l == 4 * wp and steps == -(-wp // 32) and 3 <= steps <= 5
    assert rpt == ppt * k and rpt % 4 == 0 and stages >= 2 and warps in (16, 20, 24) and (l % 2 == 0 or warps <= 20)
    seen = set()
    for p in range(l):
        blk, col = slot[p] >> 7, slot[p] & 127
        assert col < 64 and blk < (2 if l > 128 else 1)
        assert (blk, col, p & 1) not in seen
        seen.add((blk, col, p & 1))

    def word(s, i):     # the kernel's lane <-> word map
        return 32 * s + i

    def wavefronts(bank_of):
        tot = 0
        for s in range(steps):
            for j in range(4):
                c = {}
                for i in range(32):
                    if word(s, i) < wp:
                        b = bank_of((4 * word(s, i) + j) % l)
                        c[b] = c.get(b, 0) + 1
                tot += max(c.values())
        return tot
    assert wavefronts(lambda p: slot[p] & 31) <= wavefronts(lambda p: ((p & 127) >> 2) & 31)
    if l == 150:   # 4 x 150 bp: 3 of the 10 even steps cost two wavefronts, the others one
        assert wavefronts(lambda p: slot[p] & 31) == 26


def test_period_layout_rejects_other_lengths():
    import ctypes as C
    L = capi.lib()
    L.qb_period_layout.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    for l in (0, 10, 31, 161, 255, 257, 300):   # odd lengths need 4 reads per period: up to 159 bp
        assert L.qb_period_layout(l, 1, None, None) == -1


def _bgzf(data: bytes, level=6, block=65280, extra=b""):
    import struct
    import zlib
    out = []
    for o in range(0, max(len(data), 1), block):
        part = data[o:o + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = c.compress(part) + c.flush()
        xlen = 6 + len(extra)
        out.append(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff]) + struct.pack("<H", xlen) + extra + b"BC" +
                   struct.pack("<HH", 2, 12 + xlen + len(raw) + 8 - 1) + raw + struct.pack("<II", zlib.crc32(part), len(part)))
    return out


def test_bgzf_fit_walks_whole_blocks_only():
    """qb_bgzf_fit() is pure host code (block layout: reference klib/bgzf.c:63-71, trailer 261-266): how many bytes of
    a buffer are whole BGZF blocks whose text fits a slot -- what a caller of qb_bgzf_submit() cuts its chunks by."""
    import ctypes as C
    L = capi.lib()
    rng = np.random.default_rng(1)
    data = rng.integers(65, 70, size=300_000, dtype=np.uint8).tobytes()
    blocks = _bgzf(data, block=40_000, extra=b"XY\x01\x00z")
    buf = b"".join(blocks)
    whole, text = C.c_uint64(), C.c_uint64()

    def fit(b, cap):
        rc = L.qb_bgzf_fit(b, len(b), cap, C.byref(whole), C.byref(text))
        return rc, whole.value, text.value

    assert fit(buf, 1 << 30) == (0, len(buf), len(data))
    assert fit(buf[:-1], 1 << 30) == (0, len(buf) - len(blocks[-1]), len(data) - (len(data) - 40_000 * 7))   # last block cut short
    assert fit(buf, 100_000) == (0, len(blocks[0]) + len(blocks[1]), 80_000)                                # text capacity
    assert fit(buf[:10], 1 << 30) == (0, 0, 0)                                                              # less than a header
    assert fit(b"", 1 << 30) == (0, 0, 0)
    import gzip
    assert fit(gzip.compress(data), 1 << 30)[0] == -7                                                       # gzip, not BGZF
    bad = bytearray(buf)
    bad[len(blocks[0]) + 3] = 0                                                                             # FLG of block 2
    assert fit(bytes(bad), 1 << 30)[0] == -7


def test_reader_raw_mode_delivers_the_decompressed_bytes(tmp_path):
    """fqr_read_raw(): the stream without framing (what qb_text_submit takes), through every inflate back end."""
    import ctypes as C
    import gzip
    L = capi.lib()
    rng = np.random.default_rng(2)
    data = b"".join(b"@r%d\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 100)) + b"\n+\n" + b"I" * 100 + b"\n"
                    for i in range(20_000))
    files = {"plain.fq": data, "one.fq.gz": gzip.compress(data, 1),
             "members.fq.gz": b"".join(gzip.compress(data[o:o + 700_000], 1) for o in range(0, len(data), 700_000)),
             "bgzf.fq.gz": b"".join(_bgzf(data, 1)) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")}
    for name, content in files.items():
        p = tmp_path / name
        p.write_bytes(content)
        r = L.fqr_open(str(p).encode())
        assert r
        got = bytearray()
        buf = (C.c_uint8 * 300_001)()
        while True:
            n = L.fqr_read_raw(r, buf, len(buf))
            got += bytes(buf[:n])
            if n < len(buf):
                break
        assert L.fqr_status(r) == -1 and bytes(got) == data, name
        L.fqr_close(r)
