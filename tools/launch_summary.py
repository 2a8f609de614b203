#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean time, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, gi, bi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size'), H.index('Block Size')
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        d[r[ki][:70]].append((float(r[vi].replace(',', '')), r[gi], r[bi]))
tot = sum(x[0] for v in d.values() for x in v)
for k, v in sorted(d.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    s = sum(x[0] for x in v)
    print(f"{k:70s} n={len(v):3d} mean={s / len(v) / 1e3:9.1f} us share={100 * s / tot:5.1f}%  grid={v[-1][1]} block={v[-1][2]}")
