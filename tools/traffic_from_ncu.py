#!/usr/bin/env python3
"""profiles/traffic.json from one `ncu --set full` report: DRAM bytes read + written of ONE launch of
the dominant kernel (the last one captured; every captured launch is listed so the file can be
checked against the summary next to it).  bench.py reports the total as roofline.traffic."""
import csv
import json
import subprocess
import sys

rep, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
    return [float(r[i].replace(",", "")) * scale for r in rows[2:]]


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
dur = [float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) for r in rows[2:]]
launches = [{"kernel": r[hdr.index("Kernel Name")][:60], "dram_read": int(a), "dram_write": int(b),
             "duration": d, "duration_unit": units[hdr.index("gpu__time_duration.sum")]} for r, a, b, d in zip(rows[2:], rd, wr, dur)]
last = launches[-1]
json.dump({"source": source, "launch_used": len(launches) - 1, "launches_captured": launches,
           "count_stream_kernel_dram_bytes_per_launch": last["dram_read"] + last["dram_write"],
           "count_stream_kernel_dram_bytes_read": last["dram_read"], "count_stream_kernel_dram_bytes_write": last["dram_write"],
           "algorithmic_bytes_per_launch": 3800000000}, open(out, "w"), indent=1)
print(open(out).read())
