#!/usr/bin/env python3
"""Groups the SASS of one ncu report into runs of equal execution count and prints, per run,
the executed warp-instructions, the share of stall samples and the mean active lanes."""
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[which]
h = rows[hi]
si, ni, ii, ti = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[0] == "Address":
        break
    try:
        data.append((int(r[0], 16), r[si], int(r[ni] or 0), int(r[ii] or 0), int(r[ti] or 0)))
    except ValueError:
        pass
base = data[0][0]
tot = sum(d[3] for d in data)
ts = sum(d[2] for d in data)
print("total warp-instructions", tot, "samples", ts)
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][3] - data[start][3]) > 0.02 * max(data[start][3], 1):
        ex = sum(d[3] for d in data[start:i])
        sm = sum(d[2] for d in data[start:i])
        th = sum(d[4] for d in data[start:i])
        if ex > 0.004 * tot:
            print(f"off {data[start][0]-base:#06x}-{data[i-1][0]-base:#06x} n={i-start:4d} exec/instr={data[start][3]:9d} "
                  f"sum={ex/1e6:8.1f}M ({100*ex/tot:4.1f}%) samples={100*sm/ts:4.1f}% lanes={th/max(ex,1):.1f}")
        start = i
