#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: SASS instructions sorted by executed count / samples."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
mode = sys.argv[3] if len(sys.argv) > 3 else "linear"
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(f"total warp instructions {tot_inst}, samples {tot_samp}, sass lines {len(data)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def fmt(i, r):
    ie = int(r[ix["Instructions Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    sts = " ".join(f"{n}:{c}" for c, n in st if c)
    return f"{i:5d} {100*ie/tot_inst:5.2f}% {100*sm/max(tot_samp,1):5.2f}% thr={r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']][:70]:70s} {sts}"


if mode == "linear":
    for i, r in enumerate(data):
        if int(r[ix["Instructions Executed"]] or 0) * 2000 >= tot_inst or int(r[ix["# Samples"]] or 0) * 300 >= tot_samp:
            print(fmt(i, r))
else:
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
    for i in order:
        print(fmt(i, data[i]))
