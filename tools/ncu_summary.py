#!/usr/bin/env python3
"""Prints the handful of ncu raw-page metrics we track, plus the top stall sites, for one report."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "smsp__average_warp_latency",
        "smsp__average_warps_issue_stalled", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_bytes.sum ", "dram__throughput",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "lts__t_sectors.sum",
        "lts__t_sectors_lookup_miss.sum", "lts__t_sectors_lookup_hit.sum", "sm__inst_executed_pipe",
        "lts__average_t_sector", "sm__cycles_active.avg"]
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:60])
    for i, h in enumerate(hdr):
        if any(h.startswith(w.strip()) for w in want):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "issue_stalled" in h and v < 0.3:
                continue
            print(f"  {h:78s} {units[i]:10s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
si, ni, ii = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[0] == "Address":
        break
    try:
        data.append((r[si], int(r[ni] or 0), int(r[ii] or 0), r))
    except ValueError:
        pass
ts, ti = sum(d[1] for d in data), sum(d[2] for d in data)
print(f"SASS lines {len(data)}  samples {ts}  warp-instructions {ti}")
for idx, (s, n, i, r) in sorted(enumerate(data), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    top = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
    print(f"#{idx:4d} {n / ts * 100:5.1f}%  exec={i:10d}  {s[:70]:70s} {top}")
