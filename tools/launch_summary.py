#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import collections
import csv
import io
import sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in rows:
    k = r["Kernel Name"].split("(")[0]
    ms = float(r["Metric Value"]) / 1e6
    agg[k][0] += 1
    agg[k][1] += ms
    seq.append((k, ms))
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.3f ms (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%4d %10.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
if len(sys.argv) > 2:
    n = int(sys.argv[2])
    print("last %d launches:" % n)
    for k, ms in seq[-n:]:
        print("   %-40s %9.3f ms" % (k, ms))
