"""Opt-in side outputs (named by north_star, NOT computed by the reference -- no reference
oracle, "parity unpinned"): per-position N counts, per-position quality sums, per-read mean-quality distribution.
The CUDA path is checked against the scalar CPU restatement oracle/quack_oracle.c:qo_extras() and the main result
must stay what the reference computes (N folded into A) whether or not the extras pass runs."""
import json
import os
import subprocess

import numpy as np
import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, capi
from quack_b200.build import quack_bin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


def _batch(seed, n, lmin, lmax):
    # 5 % N / n, qualities over the whole printable range (clamping and the > 123 corner included)
    seq, qual, off, ln = util.random_batch(seed, n, lmin, lmax, alphabet=b"ACGTNn", probs=(.24, .24, .24, .23, .03, .02),
                                           qlo=0, qhi=93, plant=0.1)
    return seq, qual, off, ln


@pytest.mark.parametrize("shape,kernel", [((150, 150), capi.KERNEL_AUTO), ((35, 300), capi.KERNEL_AUTO), ((100, 100), capi.KERNEL_FUSED),
                                          ((1, 40), capi.KERNEL_AUTO), ((20, 700), capi.KERNEL_AUTO)],
                         ids=["period", "flat", "fused", "tiny", "long"])
def test_extras_equal_the_cpu_restatement(shape, kernel):
    table = util.oracle_table()
    batches = [_batch(50 + i, 6000, *shape) for i in range(3)]
    cap = max(shape[1], 64)
    with capi.Context(cap, n_mates=2, adapter_keys=table.keys(), batch_bytes=1 << 20, ring_depth=3, kernel=kernel) as ctx:
        ctx.extras_enable()
        for i, b in enumerate(batches):
            ctx.accumulate_host(i % 2, *b)
        for mate in (0, 1):
            mine = [b for i, b in enumerate(batches) if i % 2 == mate]
            want_n, want_q, want_m = (np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint64), np.zeros(94, dtype=np.uint64))
            for b in mine:
                n, q, m = po.extras(*b, rows=cap)
                want_n += n
                want_q += q
                want_m += m
            got_n, got_q, got_m = ctx.extras_finish(mate)
            ml = len(got_n)
            assert np.array_equal(got_n, want_n[:ml]) and not want_n[ml:].any(), (shape, mate)
            assert np.array_equal(got_q, want_q[:ml]), (shape, mate)
            assert np.array_equal(got_m, want_m) and int(got_m.sum()) == sum(len(b[2]) for b in mine)
            # the main result is the reference's whether or not the extras pass ran
            res = ctx.finish(mate)
            want = po.accumulate_batch(*mine[0], table)
            for b in mine[1:]:
                w = po.accumulate_batch(*b, table)
                rows = np.zeros((max(want.max_length, w.max_length), capi.ROW), dtype=np.uint64)
                rows[: want.max_length] += want.rows
                rows[: w.max_length] += w.rows
                want = capi.Result(rows, max(want.max_length, w.max_length), want.n_reads + w.n_reads)
            util.assert_same(res, want, f"main result with extras on, {shape}")


def test_extras_on_the_text_path_and_late_enable():
    seq, qual, off, ln = _batch(7, 4000, 60, 160)
    qual = np.clip(qual, 35, 120).astype(np.uint8)      # (FASTQ text: keep '@' '+' out of first columns, no control bytes)
    text = b"".join(b"@r\n" + seq[o:o + l].tobytes() + b"\n+\n" + qual[o:o + l].tobytes() + b"\n" for o, l in zip(off, ln))
    with capi.Context(160, batch_bytes=1 << 20, ring_depth=3) as ctx:
        ctx.extras_enable()
        ctx.text_accumulate(0, text, [len(text) // 3, len(text) // 2])
        assert ctx.text_status(0) == (4000, 0)
        got_n, got_q, got_m = ctx.extras_finish(0)
        n, q, m = po.extras(seq, qual, off, ln, rows=160)
        ml = len(got_n)
        assert np.array_equal(got_n, n[:ml]) and np.array_equal(got_q, q[:ml]) and np.array_equal(got_m, m)
    with capi.Context(160, batch_bytes=1 << 20, ring_depth=2) as ctx:
        ctx.accumulate_host(0, seq, qual, off, ln)
        with pytest.raises(capi.QbError):
            ctx.extras_enable()
        with pytest.raises(capi.QbError):
            ctx.extras_finish(0)


def test_cli_extras_json_leaves_the_svg_alone(tmp_path):
    seq, qual, off, ln = _batch(9, 3000, 80, 151)
    qual = np.clip(qual, 35, 120).astype(np.uint8)
    p = tmp_path / "x.fq"
    p.write_bytes(b"".join(b"@r\n" + seq[o:o + l].tobytes() + b"\n+\n" + qual[o:o + l].tobytes() + b"\n" for o, l in zip(off, ln)))
    js = str(tmp_path / "extras.json")
    a = subprocess.run([quack_bin(), "-u", str(p), "-a", util.ADAPTER_FA], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    b = subprocess.run([quack_bin(), "-u", str(p), "-a", util.ADAPTER_FA], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=dict(os.environ, QB_EXTRAS_JSON=js))
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout
    ex = json.load(open(js))["mates"][0]
    n, q, m = po.extras(seq, qual, off, ln, rows=151)
    assert ex["n_count"] == n.tolist() and ex["qual_sum"] == q.tolist() and ex["mean_quality_hist"] == m.tolist()
