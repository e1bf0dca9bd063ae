#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    k = r[h.index("Kernel Name")][:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[h.index("Metric Value")])
tot = sum(a[1] for a in agg.values())
print(f"{'launches':>8} {'total us':>12} {'share':>6}  kernel")
for k, a in agg.items():
    print(f"{a[0]:8d} {a[1] / 1e3:12.1f} {100 * a[1] / tot:5.1f}%  {k}")
