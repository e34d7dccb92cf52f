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
