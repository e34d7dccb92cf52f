#!/usr/bin/env python
"""Segment summary of an `ncu --page source --csv` dump (any header offset): runs of SASS lines with the same
execution count -> share of warp instructions / stall samples and warp instructions per read.
usage: ncu_segments2.py source.csv [n_reads]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_reads = float(sys.argv[2]) if len(sys.argv) > 2 else 2e6
h = 0
while 'Source' not in rows[h]:
    h += 1
ix = {x: i for i, x in enumerate(rows[h])}
data = rows[h + 1:]


def ie(i):
    return int(data[i][ix['Instructions Executed']] or 0)


def sm(i):
    return int(data[i][ix['# Samples']] or 0)


tot = sum(ie(i) for i in range(len(data)))
ts = sum(sm(i) for i in range(len(data)))
print('total warp instr', tot, 'per read %.2f' % (tot / n_reads), 'samples', ts)
seg, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(ie(i) - ie(i - 1)) > 0.1 * max(ie(i), ie(i - 1), 1):
        seg.append((start, i - 1))
        start = i
for a, b in seg:
    s = sum(ie(i) for i in range(a, b + 1))
    p = sum(sm(i) for i in range(a, b + 1))
    if s * 200 > tot or p * 100 > ts:
        ops = {}
        for i in range(a, b + 1):
            op = [o for o in data[i][ix['Source']].split() if not o.startswith('@')]
            op = op[0].split('.')[0] if op else '?'
            ops[op] = ops.get(op, 0) + 1
        top = ' '.join(f'{k}:{v}' for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
        print(f'{a:5d}-{b:5d} n={b - a + 1:4d} x{ie(a):9d} instr={100 * s / tot:5.1f}% ({s / n_reads:6.2f}/read) '
              f'samples={100 * p / ts:5.1f}%  {top}')
