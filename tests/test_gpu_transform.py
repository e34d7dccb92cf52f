"""Device-side transform() (reference quack.c:230-293): qb_finish_transformed() must return
what the reference's transform() makes of the raw accumulator -- binning of reads longer than 3000 bp with its in-place
quirks (bin b keeps position b's kmer_count, the last bin is dropped), running kmer_count, integer score percentages,
single-precision ceil for the length / kmer columns.  Oracle: qo_transform (pinned against the reference's transform()
by tests/test_oracle.py and the KAT-F goldens)."""
import numpy as np
import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


@pytest.mark.parametrize("lmin,lmax,n", [(150, 150, 5000), (35, 300, 5000), (2990, 3000, 40), (2995, 3001, 40), (3001, 3100, 40),
                                         (100, 9999, 60), (20000, 45000, 12), (99_900, 100_101, 4)],
                         ids=["150", "ragged", "3000", "3001", "3100", "10k", "45k", "100k"])
@pytest.mark.parametrize("adapters", [False, True], ids=["noad", "ad"])
def test_device_transform_equals_the_oracle(lmin, lmax, n, adapters):
    table = util.oracle_table() if adapters else None
    batch = util.random_batch(lmax + n, n, lmin, lmax, plant=0.3)
    with capi.Context(max(lmax, 64), adapter_keys=table.keys() if adapters else None, batch_bytes=4 << 20, ring_depth=2) as ctx:
        ctx.accumulate_host(0, *batch)
        raw = ctx.finish(0)
        rows, ml, nr, orig = ctx.finish_transformed(0)
    want = po.accumulate_batch(*batch, table)
    util.assert_same(raw, want, "raw rows")
    wrows, wml, worig = po.transform(want.rows, want.max_length, want.n_reads)
    assert (ml, nr, orig) == (wml, want.n_reads, worig)
    if not np.array_equal(rows, wrows[:wml]):
        bad = np.argwhere(rows != wrows[:wml])
        p, c = bad[0]
        raise AssertionError(f"{len(bad)} cells differ; first at row {p} col {c}: got {rows[p, c]} want {wrows[p, c]}")


def test_cli_device_transform_same_svg(tmp_path):
    """The `quack` program with QB_DEVICE_TRANSFORM=1: same SVG, for short reads and for reads that get binned."""
    import os
    import subprocess
    from quack_b200.build import quack_bin
    for name, (lmin, lmax, n) in {"short": (100, 151, 3000), "long": (3000, 12000, 60)}.items():
        seq, qual, off, ln = util.random_batch(len(name), n, lmin, lmax, qlo=2, qhi=40)
        p = tmp_path / (name + ".fq")
        p.write_bytes(b"".join(b"@r\n" + seq[o:o + l].tobytes() + b"\n+\n" + qual[o:o + l].tobytes() + b"\n" for o, l in zip(off, ln)))
        outs = []
        for env in ({}, {"QB_DEVICE_TRANSFORM": "1"}):
            for args in (["-u", str(p)], ["-u", str(p), "-a", util.ADAPTER_FA]):
                r = subprocess.run([quack_bin(), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env))
                assert r.returncode == 0, r.stderr
                outs.append((r.stdout, b"Binning" in r.stderr))
        assert outs[0] == outs[2] and outs[1] == outs[3], name
        assert outs[0][1] == (name == "long")
