#!/usr/bin/env python
"""Host decode throughput: inflate + framing + packing of one synthetic
150-bp FASTQ file, no GPU involved.  gzip file (members of 16 MiB of text, written by quack_b200/bin/qb_gen_fastq)
through gzread() (the reference's path, quack.c:187) against the member pool at 2..N threads, and the same reads as
a BGZF file through the block pool.  One JSON line per run.   usage: decode_bench.py [n_reads] [max_threads]"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quack_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
tmax = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
with tempfile.TemporaryDirectory() as d:
    gz, bg = os.path.join(d, "s.fq.gz"), os.path.join(d, "s.fq.bgz")
    import subprocess
    from quack_b200.build import gen_bin
    for path, mode in ((gz, "gz"), (bg, "bgzf")):
        subprocess.run([gen_bin(), path, "2", "1", "0", str(n), "150", "150", "0.1", mode, "1"], check=True)
    text = n * (14 + 150 + 3 + 150 + 1)
    ts = [t for t in (2, 4, 8, 12, 16, 24, 32) if t <= tmax]
    runs = [("gzip/gzread", gz, 1)] + [("gzip/member pool", gz, t) for t in ts] + [("bgzf/pool", bg, t) for t in ts]
    base = None
    for name, path, t in runs:
        best = None
        for _ in range(2):
            r = capi.decode_throughput(path, t)
            if best is None or r["seconds"] < best["seconds"]:
                best = r
        assert best["reads"] == n and best["status"] == -1
        rate = best["reads"] / best["seconds"]
        base = base or rate
        print(json.dumps({"input": name, "file_MB": round(os.path.getsize(path) / 1e6, 1), "threads": best["threads"],
                          "reads": n, "seconds": round(best["seconds"], 3), "Mreads_s": round(rate / 1e6, 3),
                          "text_MBps": round(best["text_bytes"] / best["seconds"] / 1e6, 1),
                          "wait_for_inflate_s": round(best["wait_inflate_s"], 3), "vs_gzread": round(rate / base, 2),
                          "cores": os.cpu_count()}), flush=True)
