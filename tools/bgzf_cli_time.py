#!/usr/bin/env python3
"""The `sgcount` CLI on BGZF input, device ingest, on 1 GPU and on all visible GPUs:
   (a) eight samples of `n_reads` reads each (one device per sample), and
   (b) ONE sample of `big_reads` reads whose blocks the host cuts into waves at record boundaries and
       deals to the devices (config 5 end to end from gzip).
Identical tables required; prints the phase times.    usage: bgzf_cli_time.py [n_reads [big_reads]]"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sgcount_b200 import synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 20
big_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 64 << 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "sgcount_b200", "lib", "sgcount")
seed = 0xB2000004
arr = synth.make_library(seed, 200000, 20)
tmp = tempfile.mkdtemp(prefix="sgc_bgzf_")
lib = os.path.join(tmp, "lib.fa")
open(lib, "wb").write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
truth = [(False, 7), (True, 30), (False, 0), (True, 12), (False, 23), (True, 5), (False, 40), (True, 0)]
paths = []
for s, (rev, off) in enumerate(truth):
    p = os.path.join(tmp, f"b{s}.fastq.gz")
    synth.Sample(seed, s, arr, 75, off, rev).write_fastq_bgzf(p, 0, n_reads, gz_level=1)
    paths.append(p)
big = os.path.join(tmp, "big.fastq.gz")
t0 = time.time()
synth.Sample(seed, 100, arr, 75, 9, False).write_fastq_bgzf(big, 0, big_reads, gz_level=1)
print(f"big sample: {big_reads} reads, {os.path.getsize(big) / 1e6:.0f} MB of BGZF, written in {time.time() - t0:.1f} s", flush=True)
n_dev = torch.cuda.device_count()


def run(args, label):
    best = None
    for _ in range(2):
        p = subprocess.run([exe, "-l", lib, *args, "--timing"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
        best = t if best is None or t["count_s"] < best["count_s"] else best
    print(f"{label}: count_s {best['count_s']:.3f}  {best['reads'] / best['count_s'] / 1e6:.1f} M reads/s  device samples "
          f"{best['device_ingest_samples']}  shards {best['read_shards_per_sample']}  blocks {best['device_blocks']}  "
          f"phases index/create/waves/finish {best['device_phases_s']}  tables {best['device_tables_s']:.2f} s", flush=True)
    return best


tables = {}
for gpus in sorted({1, n_dev}):
    out = os.path.join(tmp, f"eight{gpus}.tsv")
    run(["-i", *paths, "-o", out, "-t", str(max(gpus, 2)), "--gpus", str(gpus)], f"eight samples x {n_reads} reads, gpus={gpus}")
    tables[gpus] = open(out, "rb").read()
assert len(set(tables.values())) == 1
tables = {}
for gpus in sorted({1, n_dev}):
    out = os.path.join(tmp, f"big{gpus}.tsv")
    run(["-i", big, "-o", out, "--gpus", str(gpus), "--read-shards", str(gpus)], f"one sample of {big_reads} reads, gpus={gpus}")
    tables[gpus] = open(out, "rb").read()
assert len(set(tables.values())) == 1
print("tables identical on 1 and", n_dev, "GPUs")
