"""Shared helpers for the parity tests: seeded synthetic batches in the C-ABI layout."""
import os

import numpy as np

from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ADAPTER_FA = os.path.join(GOLDEN, "adapters_all.fa")


def adapter_records():
    recs, cur = [], None
    for line in open(ADAPTER_FA, "rb").read().split(b"\n"):
        if line.startswith(b">"):
            if cur is not None:
                recs.append(cur)
            cur = b""
        elif cur is not None:
            cur += line.strip()
    if cur:
        recs.append(cur)
    return recs


def oracle_table():
    return po.AdapterTable.from_file(ADAPTER_FA)


def pack(reads):
    """[(seq bytes, qual bytes)] -> (seq, qual, offset, length) arrays, concatenated and unpadded."""
    lens = np.array([len(s) for s, _ in reads], dtype=np.uint32)
    off = np.zeros(len(reads), dtype=np.uint32)
    if len(reads) > 1:
        off[1:] = np.cumsum(lens[:-1], dtype=np.uint64).astype(np.uint32)
    seq = np.frombuffer(b"".join(s for s, _ in reads), dtype=np.uint8)
    qual = np.frombuffer(b"".join(q for _, q in reads), dtype=np.uint8)
    return seq, qual, off, lens


def random_batch(seed, n, lmin, lmax, alphabet=b"ACGTN", probs=(.2495, .2495, .2495, .2495, .002),
                 qlo=2, qhi=41, plant=0.0, plant_seq=b"AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"):
    """Vectorised random reads; `plant` = fraction of reads that switch to an adapter at a random position."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lmin, lmax + 1, size=n).astype(np.uint32)
    off = np.zeros(n, dtype=np.uint32)
    off[1:] = np.cumsum(lens[:-1], dtype=np.uint64).astype(np.uint32)
    total = int(lens.sum())
    seq = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=total, p=probs)
    qual = (rng.integers(qlo, qhi + 1, size=total) + 33).astype(np.uint8)
    if plant > 0:
        ad = np.frombuffer(plant_seq, dtype=np.uint8)
        for r in np.flatnonzero(rng.random(n) < plant):
            l = int(lens[r])
            if l < 2:
                continue
            at = int(rng.integers(0, l))
            m = min(len(ad), l - at)
            seq[off[r] + at: off[r] + at + m] = ad[:m]
    return seq, qual, off, lens


def assert_same(got, want, what=""):
    assert (got.max_length, got.n_reads) == (want.max_length, want.n_reads), what
    if not np.array_equal(got.rows, want.rows):
        bad = np.argwhere(got.rows != want.rows)
        p, c = bad[0]
        raise AssertionError(f"{what}: {len(bad)} cells differ; first at pos {p} col {c}: "
                             f"got {got.rows[p, c]} want {want.rows[p, c]}")
