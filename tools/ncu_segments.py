#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump by code segment (runs of SASS lines with the same
execution count): share of warp instructions, of stall samples, and instructions per 32 bases.
usage: ncu_segments.py source.csv [total_bases]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
bases = float(sys.argv[2]) if len(sys.argv) > 2 else 300e6
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def ie(i):
    return int(data[i][ix["Instructions Executed"]] or 0)


def sm(i):
    return int(data[i][ix["# Samples"]] or 0)


tot = sum(ie(i) for i in range(len(data)))
ts = sum(sm(i) for i in range(len(data)))
print(f"kernel {rows[0][1][:60]}  warp-instr {tot}  = {tot / (bases / 32):.2f} per 32 bases; samples {ts}")
seg, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(ie(i) - ie(i - 1)) > 0.15 * max(ie(i), ie(i - 1), 1):
        seg.append((start, i - 1))
        start = i
for a, b in seg:
    s = sum(ie(i) for i in range(a, b + 1))
    p = sum(sm(i) for i in range(a, b + 1))
    if s * 250 > tot or p * 100 > ts:
        ops = {}
        for i in range(a, b + 1):
            op = data[i][ix["Source"]].split()
            op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"{a:5d}-{b:5d} n={b - a + 1:4d} x{ie(a):>9d} instr={100 * s / tot:6.2f}% ({s / (bases / 32):5.2f}/32b) "
              f"samples={100 * p / max(ts, 1):6.2f}%  {top}")
