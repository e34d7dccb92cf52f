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
