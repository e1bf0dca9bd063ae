#!/usr/bin/env python3
"""A/B of two builds of the `sgcount` CLI on the SAME box (boxes of the pool differ by more than the
effects looked for): eight BGZF samples on one GPU, runs interleaved.   usage: cli_ab.py exeA exeB [n_reads [rounds]]"""
import json
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgcount_b200 import synth

exes = sys.argv[1:3]
n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 4 << 20
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
seed = 0xB2000004
arr = synth.make_library(seed, 200000, 20)
tmp = tempfile.mkdtemp(prefix="sgc_ab_")
lib = os.path.join(tmp, "lib.fa")
open(lib, "wb").write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
truth = [(False, 7), (True, 30), (False, 0), (True, 12), (False, 23), (True, 5), (False, 40), (True, 0)]
paths = []
for s, (rev, off) in enumerate(truth):
    p = os.path.join(tmp, f"b{s}.fastq.gz")
    synth.Sample(seed, s, arr, 75, off, rev).write_fastq_bgzf(p, 0, n_reads, gz_level=1)
    paths.append(p)
times = {e: [] for e in exes}
tables = set()
for r in range(rounds):
    for e in exes:
        out = os.path.join(tmp, "out.tsv")
        p = subprocess.run([e, "-l", lib, "-i", *paths, "-o", out, "--timing"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
        times[e].append(t["count_s"])
        tables.add(open(out, "rb").read())
        print(os.path.basename(e), f"count_s {t['count_s']:.3f} workers {t['sample_workers']} phases {t['device_phases_s']} tables {t['device_tables_s']:.2f}", flush=True)
assert len(tables) == 1
for e in exes:
    ts = sorted(times[e])
    print(f"{os.path.basename(e)}: best {ts[0]:.3f} s  median {ts[len(ts) // 2]:.3f} s  = {8 * n_reads / ts[0] / 1e6:.0f} M reads/s best")
