#!/usr/bin/env python
"""File-to-SVG wall time of the drop-in on a config-2 shaped sample (paired 150 bp, -a all.fa.gz): the
unmodified reference binary (oracle/_ref/quack, single-threaded, gzread) against the `quack` host program
of this repo on the same gzip files, and on the same reads as BGZF files (inflate pool).  Checks that all
SVGs are byte-identical.  One JSON line per run.   usage: cli_bench.py [pairs] [threads list]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quack_b200 import synth
from quack_b200.build import quack_bin

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "quack")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
threads = [int(t) for t in (sys.argv[2] if len(sys.argv) > 2 else "1,4,8").split(",")]


def run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    return time.perf_counter() - t0, r


with tempfile.TemporaryDirectory() as d:
    g = [os.path.join(d, f"s_{m}.fq.gz") for m in (1, 2)]
    b = [os.path.join(d, f"b_{m}.fq.gz") for m in (1, 2)]
    for m in (1, 2):
        synth.write_fastq(g[m - 1], 2, m, pairs, 150, 0.1, gz_level=1)
        synth.write_fastq(b[m - 1], 2, m, pairs, 150, 0.1, gz_level=1, bgzf=True)
    common = ["-a", synth.ADAPTER_FA, "-n", "cfg2 sample"]
    svg = None
    if os.path.exists(REF_BIN):
        dt, r = run([REF_BIN, "-1", g[0], "-2", g[1], *common])
        svg = r.stdout
        print(json.dumps({"program": "reference quack (1 thread, gzread)", "input": "gzip", "pairs": pairs,
                          "seconds": round(dt, 3), "Mreads_s": round(2 * pairs / dt / 1e6, 3)}), flush=True)
    runs = [("gzip", g, 1)] + [("bgzf", b, t) for t in threads]
    for name, files, t in runs:
        js = os.path.join(d, "stats.json")
        best = None
        for _ in range(2):
            dt, r = run([quack_bin(), "-1", files[0], "-2", files[1], *common],
                        {"QUACK_DECODE_THREADS": str(t), "QB_STATS_JSON": js})
            assert r.returncode == 0, r.stderr
            if best is None or dt < best[0]:
                best = (dt, r, json.load(open(js)))
        dt, r, st = best
        if svg is None:
            svg = r.stdout
        assert r.stdout == svg, "SVG differs"
        print(json.dumps({"program": "quack_b200 quack (1 GPU)", "input": name, "decode_threads_per_file": st["decode_threads"],
                          "pairs": pairs, "seconds": round(dt, 3), "Mreads_s": round(2 * pairs / dt / 1e6, 3),
                          "create_s": round(st["create_s"], 3), "decode_to_counts_s": round(st["stream_s"] - st["create_s"], 3),
                          "total_s_in_process": round(st["total_s"], 3),
                          "svg_identical": True}), flush=True)
