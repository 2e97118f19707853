#!/usr/bin/env python
"""ncu source-page digest: python tools/ncu_src.py <rep> <kernel regex> [top N]"""
import csv, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ia, isrc, ism, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [(int(r[ia], 16), r[isrc].strip(), int(r[ism]), int(r[iex])) for r in rows[2:] if len(r) > iex]
base = data[0][0]
tot, totex = sum(d[2] for d in data), sum(d[3] for d in data)
print(rows[0][1][:100]); print("total samples", tot, "total warp inst", totex)
ops, smp = Counter(), Counter()
for a, s, m, e in data:
    t = s.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += e; smp[op] += m
for op, n in ops.most_common(22): print(f"  {op:10s} {n:11d} {100*n/totex:5.1f}%  samples {smp[op]}")
print("hot lines:")
for d in sorted(sorted(data, key=lambda d: -d[2])[:topn], key=lambda d: d[0]): print(f"  {d[0]-base:#7x} {d[2]:6d} {d[3]:9d} {d[1][:100]}")
