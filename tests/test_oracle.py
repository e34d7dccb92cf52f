"""Pins the parity oracle (oracle/quack_oracle.c): against the committed golden vectors that
the unmodified reference produced (tests/golden/make_golden.py), against the known answers of
SURVEY.md Appendix B, and -- when oracle/_ref/ was built -- against the reference itself on
fresh random inputs."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po

FIXTURES = {"kat_t": "kat_t.fq", "kat_k": "kat_k.fq", "framing": "framing.fq", "rand_small": "rand_small.fq.gz"}
PARSE_FIXTURES = dict(FIXTURES, kat_p="kat_p.dat")   # KAT-P: klib/test/kseq_test.dat, the reference's parser fixture


@pytest.fixture(scope="module")
def golden_raw(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_raw.npz"))


@pytest.fixture(scope="module")
def table(adapters_fa):
    return po.AdapterTable.from_file(adapters_fa)


@pytest.mark.parametrize("name", sorted(FIXTURES))
@pytest.mark.parametrize("tag", ["noad", "ad"])
def test_oracle_matches_reference_golden(name, tag, golden_raw, golden_dir, table):
    res = po.read_fastq(os.path.join(golden_dir, FIXTURES[name]), table if tag == "ad" else None)
    ml, n = golden_raw[f"{name}.{tag}.meta"]
    assert (res.max_length, res.n_reads) == (int(ml), int(n))
    assert np.array_equal(res.rows, golden_raw[f"{name}.{tag}.rows"])


def test_adapter_keys_match_reference(table, golden_dir):
    keys = table.keys()
    gold = np.load(os.path.join(golden_dir, "golden_adapter_keys.npy"))
    assert len(keys) == 333
    assert np.array_equal(keys, gold)


def _key(s):
    k = 0
    for ch in s:
        k = (k << 2) | {"A": 0, "T": 1, "C": 2, "G": 3}[ch]
    return k


def test_adapter_set_quirks(table):
    # the window starting at base 0 of a record is never inserted
    t = table.as_bytes()
    assert t[_key("AGATCGGAAG")] == 0 and t[_key("GATCGGAAGA")] == 1
    assert t[_key("TTTTTTTTTT")] == 0 and t[_key("TTTTTTTTTA")] == 1
    short = po.AdapterTable.from_records([b"ACGTACGTAC"])  # l == 10 inserts nothing
    assert len(short.keys()) == 0
    one = po.AdapterTable.from_records([b"ACGTACGTACG"])  # l == 11 inserts exactly the 2nd window
    assert list(one.keys()) == [_key("CGTACGTACG")]
    n_as_a = po.AdapterTable.from_records([b"NNNNNNNNNNNN"])  # N counts as A
    assert list(n_as_a.keys()) == [0]


def test_kat_t_known_answers(golden_dir, table):
    # SURVEY.md Appendix B, KAT-T
    r = po.read_fastq(os.path.join(golden_dir, "kat_t.fq"), None)
    assert (r.max_length, r.n_reads) == (22, 2)
    assert r.rows[0, 91] == 2                       # pos 0: A,A
    assert r.rows[4, 91] == 1 and r.rows[4, 93] == 1  # pos 4: N (counted as A) and C
    assert r.rows[10, 91] == 2
    assert all(r.rows[p, 40] == 2 for p in range(10))
    assert all(r.rows[p, 0] == 1 and r.rows[p, 40] == 1 for p in range(10, 15))
    assert all(r.rows[p, 20] == 1 and r.rows[p, 40] == 1 for p in range(15, 22))
    assert r.rows[21, 95] == 2
    assert r.rows[10, 96] == 2 and r.rows[:, 96].sum() == 2   # no -a quirk: kmer_count[10] per read
    ra = po.read_fastq(os.path.join(golden_dir, "kat_t.fq"), table)
    assert ra.rows[11, 96] == 1 and ra.rows[:, 96].sum() == 1


def test_kat_k_known_answers(golden_dir, table):
    r = po.read_fastq(os.path.join(golden_dir, "kat_k.fq"), table)
    assert (r.max_length, r.n_reads) == (31, 8)
    assert {int(p): int(r.rows[p, 96]) for p in np.flatnonzero(r.rows[:, 96])} == {10: 2, 15: 1, 16: 1}
    assert {int(p): int(r.rows[p, 95]) for p in np.flatnonzero(r.rows[:, 95])} == {4: 1, 9: 1, 10: 1, 19: 1, 21: 2, 30: 2}


def test_framing_matches_reference_reader(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "golden_parse.json")))
    for name in ("framing", "kat_k", "kat_p"):
        recs, rc = po.parse_records(os.path.join(golden_dir, PARSE_FIXTURES[name]))
        assert rc == gold[name]["rc"]
        assert [[s.decode("latin1"), q.decode("latin1") if q else None] for s, q in recs] == gold[name]["records"]
    recs, rc = po.parse_records(os.path.join(golden_dir, "framing.fq"))
    assert rc == -2 and len(recs) == 3   # truncated 4th record ends the stream (kseq.h:216)
    assert recs[0] == (b"ACGTACGTACGTACGTACGTAA", b"IIIIIIIIII@IIIIIIIIIII")
    assert recs[1] == (b"acgtnACGTN", b"!!!!!55555")


def test_transform_kat_f(golden_dir):
    for kat in json.load(open(os.path.join(golden_dir, "golden_transform.json"))):
        rows = np.zeros((2, po.ROW), dtype=np.uint64)
        rows[0, 95] = rows[0, 96] = kat["count"]
        rows[0, 40], rows[0, 2] = 7, 1
        t, ml, orig = po.transform(rows, 2, kat["n_reads"])
        assert (int(t[0, 95]), int(t[0, 96]), int(t[1, 96])) == (kat["length_pct"], kat["kmer_pct"], kat["kmer_pct_row1"])
        assert (int(t[0, 40]), int(t[0, 2])) == (87, 12) == (kat["score40_pct"], kat["score2_pct"])
    # the discriminating single-precision case of Appendix B
    rows = np.zeros((2, po.ROW), dtype=np.uint64)
    rows[0, 95] = 20000001
    assert int(po.transform(rows, 2, 200000000)[0][0, 95]) == 10


def test_base_code_domain():
    for ch in range(256):
        want = {"C": 2, "c": 2, "G": 3, "g": 3, "T": 1, "t": 1}.get(chr(ch), None)
        if chr(ch).isalpha() and chr(ch).upper() <= "T" and ch < 128:
            assert po.base_code(ch) == (want or 0), chr(ch)


# ------------------------------------------------------------- live reference (build container only)

needs_ref = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
def test_base_code_vs_reference_lookup():
    for ch in list(range(65, 85)) + list(range(97, 117)):
        assert po.base_code(ch) == po.ref().qref_base_code(ch)


def _write_random_fastq(path, rng, n, lmin, lmax, multiline=False):
    with open(path, "wb") as f:
        for i in range(n):
            l = int(rng.integers(lmin, lmax + 1))
            s = rng.choice(np.frombuffer(b"ACGTNacgtnRSKM", dtype=np.uint8), size=l,
                           p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .004, .004, .004, .004, .004]).tobytes()
            q = (rng.integers(0, 91, size=l).astype(np.uint8) + 33).tobytes()
            if multiline and l > 20:
                c = l // 3
                s = s[:c] + b"\n" + s[c:]
                q = q[:c] + b"\n" + q[c:]
            f.write(b"@x%d some comment\n%s\n+\n%s\n" % (i, s, q))


@needs_ref
@pytest.mark.parametrize("seed,lmin,lmax,multiline", [(1, 1, 40, False), (2, 150, 150, False), (3, 30, 320, True)])
def test_oracle_vs_live_reference_random(tmp_path, adapters_fa, table, seed, lmin, lmax, multiline):
    rng = np.random.default_rng(seed)
    p = str(tmp_path / "r.fq")
    _write_random_fastq(p, rng, 3000, lmin, lmax, multiline)
    for ad, tab in ((None, None), (adapters_fa, table)):
        want = po.ref_read_fastq(p, ad)
        got = po.read_fastq(p, tab)
        assert (got.max_length, got.n_reads) == (want.max_length, want.n_reads)
        assert np.array_equal(got.rows, want.rows)
        # transform restatement against the reference transform on real data
        t_got = po.transform(got.rows, got.max_length, got.n_reads)
        t_want = po.ref_transform(want.rows, want.max_length, want.n_reads)
        assert t_got[1:] == t_want[1:] and np.array_equal(t_got[0], t_want[0])


@needs_ref
def test_transform_binning_vs_reference():
    rng = np.random.default_rng(7)
    rows = rng.integers(0, 50, size=(3456, po.ROW)).astype(np.uint64)
    got = po.transform(rows, 3456, 1234)
    want = po.ref_transform(rows, 3456, 1234)
    assert got[1:] == want[1:] == (34, 3456)
    assert np.array_equal(got[0][:34], want[0][:34])


def test_extras_restatement_equals_its_definition():
    """qo_extras() (PARITY UNPINNED: the reference has no such arrays) against the definition
    written out with numpy, so that the checker of the CUDA extras pass is itself checked."""
    import qb_testutil as util
    seq, qual, off, ln = util.random_batch(3, 500, 1, 90, alphabet=b"ACGTNn", probs=(.2, .2, .2, .2, .1, .1), qlo=0, qhi=93)
    qual = qual.copy()
    qual[::97] = 20          # below '!' : clamped for the mean, outside the heatmap's range for the sum
    n, q, m = po.extras(seq, qual, off, ln, rows=90)
    wn, wq, wm = np.zeros(90, dtype=np.uint64), np.zeros(90, dtype=np.uint64), np.zeros(94, dtype=np.uint64)
    for o, l in zip(off.tolist(), ln.tolist()):
        s, qq = seq[o:o + l], qual[o:o + l].astype(np.int64)
        wn[:l] += ((s == ord("N")) | (s == ord("n"))).astype(np.uint64)
        ok = (qq >= 33) & (qq <= 123)
        wq[:l] += np.where(ok, qq - 33, 0).astype(np.uint64)
        wm[int((np.clip(qq, 33, 126) - 33).sum()) // l] += 1
    assert np.array_equal(n, wn) and np.array_equal(q, wq) and np.array_equal(m, wm)
