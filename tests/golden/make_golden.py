#!/usr/bin/env python
"""Regenerates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Inputs (committed): kat_t.fq, kat_k.fq, framing.fq (record-framing edge
cases), kat_p.dat (KAT-P: the reference's own parser fixture klib/test/kseq_test.dat -- 2 FASTA records and a
multi-line FASTQ record with an empty line inside the quality block), adapters_all.fa (= zcat /root/reference/all.fa.gz, the adapter set configs 2-4 name).
Outputs (committed): rand_small.fq.gz, golden_raw.npz, golden_parse.json, golden_transform.json,
golden_adapter_keys.npy, svg/*.svg.  The reference itself cannot travel to the GPU box; these
files are how its answers do.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyoracle as po  # noqa: E402

ADAPTERS = os.path.join(HERE, "adapters_all.fa")


def make_rand_small(path):
    """1000 reads: 600 x 150 bp + 400 x uniform 12..300 bp, Phred 2..41 (+ a few higher), 0.5% N,
    lower-case runs, and TruSeq read-through planted in 15% of the reads."""
    rng = np.random.default_rng(20261017)
    adapter = b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
    out = []
    for i in range(1000):
        l = 150 if i < 600 else int(rng.integers(12, 301))
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=l)
        s[rng.random(l) < 0.005] = ord("N")
        if rng.random() < 0.15 and l > 30:
            at = int(rng.integers(5, l - 5))
            ins = np.frombuffer(adapter, dtype=np.uint8)[: l - at]
            s[at: at + len(ins)] = ins
            s[at + len(ins):] = ord("A")
        if rng.random() < 0.1:
            a, b = sorted(rng.integers(0, l, size=2))
            s[a:b] |= 0x20  # lower case
        q = np.clip(np.rint(rng.normal(36 - 12 * (np.arange(l) / 300.0) ** 2, 5)), 2, 41).astype(np.uint8) + 33
        if i % 97 == 0:
            q[rng.integers(0, l)] = 33 + 60  # a rare high score
        out.append(b"@r%d/1\n%s\n+\n%s\n" % (i, s.tobytes(), q.tobytes()))
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(b"".join(out))


def main():
    if not po.have_ref():
        po.build(ref=True)
    assert po.have_ref(), "reference not built (make -C oracle ref)"
    rand = os.path.join(HERE, "rand_small.fq.gz")
    make_rand_small(rand)

    fixtures = {
        "kat_t": os.path.join(HERE, "kat_t.fq"),
        "kat_k": os.path.join(HERE, "kat_k.fq"),
        "framing": os.path.join(HERE, "framing.fq"),
        "rand_small": rand,
    }
    raw = {}
    for name, path in fixtures.items():
        for tag, ad in (("noad", None), ("ad", ADAPTERS)):
            r = po.ref_read_fastq(path, ad)
            raw[f"{name}.{tag}.rows"] = r.rows
            raw[f"{name}.{tag}.meta"] = np.array([r.max_length, r.n_reads], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "golden_raw.npz"), **raw)

    parse = {}
    fixtures_parse = dict(fixtures, kat_p=os.path.join(HERE, "kat_p.dat"))  # KAT-P = klib/test/kseq_test.dat
    for name in ("framing", "kat_k", "kat_p"):
        recs, rc = po.ref_parse_records(fixtures_parse[name])
        parse[name] = {"rc": rc, "records": [[s.decode("latin1"), q.decode("latin1") if q else None] for s, q in recs]}
    json.dump(parse, open(os.path.join(HERE, "golden_parse.json"), "w"), indent=1)

    table = po.ref_adapter_table(ADAPTERS)
    np.save(os.path.join(HERE, "golden_adapter_keys.npy"), np.flatnonzero(table).astype(np.uint32))

    # KAT-F: transform rounding, through the reference transform() itself
    kat_f = []
    for cnt, n in [(20000001, 200000000), (1, 10**7), (3, 7), (1, 3), (33554433, 5 * 10**7),
                   (123456789, 2 * 10**8), (49999999, 5 * 10**7), (16777217, 16777217)]:
        rows = np.zeros((2, po.ROW), dtype=np.uint64)
        rows[0, 95] = cnt
        rows[0, 96] = cnt
        rows[0, 40] = 7
        rows[0, 2] = 1
        t, ml, _ = po.ref_transform(rows, 2, n)
        kat_f.append({"count": cnt, "n_reads": n, "length_pct": int(t[0, 95]), "kmer_pct": int(t[0, 96]),
                      "kmer_pct_row1": int(t[1, 96]), "score40_pct": int(t[0, 40]), "score2_pct": int(t[0, 2])})
    json.dump(kat_f, open(os.path.join(HERE, "golden_transform.json"), "w"), indent=1)

    os.makedirs(os.path.join(HERE, "svg"), exist_ok=True)
    t, k, r = fixtures["kat_t"], fixtures["kat_k"], rand
    combos = {
        "u_kat_t": ["-u", t],
        "u_kat_t_ad": ["-u", t, "-a", ADAPTERS],
        "u_kat_t_ad_name": ["-u", t, "-a", ADAPTERS, "-n", "tiny"],
        "u_kat_k_ad": ["-u", k, "-a", ADAPTERS],
        "pe_kat_t_kat_k_ad_name": ["-1", t, "-2", k, "-a", ADAPTERS, "-n", "pair"],
        "u_rand": ["-u", r],
        "u_rand_ad": ["-u", r, "-a", ADAPTERS],
        "pe_rand_kat_k": ["-1", r, "-2", k],
        "pe_rand_rand_ad_name": ["--forward", r, "--reverse", r, "--adapters", ADAPTERS, "--name", "rand x2"],
    }
    md5 = {}
    for name, args in combos.items():
        svg = po.ref_svg(args)
        if len(svg) > 100_000:  # big ones are stored gzipped (deterministic: mtime=0)
            with gzip.GzipFile(os.path.join(HERE, "svg", name + ".svg.gz"), "wb", mtime=0) as f:
                f.write(svg)
        else:
            open(os.path.join(HERE, "svg", name + ".svg"), "wb").write(svg)
        md5[name] = hashlib.md5(svg).hexdigest()
    json.dump({"args": {k_: [os.path.basename(a) if os.path.exists(a) else a for a in v] for k_, v in combos.items()},
               "md5": md5}, open(os.path.join(HERE, "svg", "index.json"), "w"), indent=1)
    print(json.dumps(md5, indent=1))


if __name__ == "__main__":
    main()
