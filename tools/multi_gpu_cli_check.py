#!/usr/bin/env python3
"""Config 4 in small: CRISPRi-shaped library (200 000 guides, gene map), 8 samples with mixed
orientations and offsets, counted by the `sgcount` CLI on 1 GPU and on all visible GPUs (one
sample per GPU at a time, the reference's fan-out over samples): the two tables must be
byte-identical and the auto-detected offsets the planted ones.  Verification aid."""
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
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "sgcount_b200", "lib", "sgcount")
seed = 0xB2000004
arr = synth.make_library(seed, 200000, 20)
tmp = tempfile.mkdtemp(prefix="sgc_mgpu_")
lib, g2s = os.path.join(tmp, "lib.fa"), os.path.join(tmp, "g2s.txt")
open(lib, "wb").write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
open(g2s, "wb").write(b"".join(b"gene.%d\tlib.%d\n" % (i // 10, i) for i in range(len(arr))))
truth = [(False, 7), (True, 30), (False, 0), (True, 12), (False, 23), (True, 5), (False, 40), (True, 0)]
paths = []
for s, (rev, off) in enumerate(truth):
    p = os.path.join(tmp, f"s{s}.fastq.gz")
    synth.Sample(seed, s, arr, 75, off, rev).write_fastq(p, 0, n_reads, reads_per_member=1 << 20, gz_level=1)
    paths.append(p)
want = "Calculated Offsets: [" + ", ".join(f"{'Reverse' if r else 'Forward'}({o})" for r, o in truth) + "]"
tables = {}
for gpus in sorted({1, torch.cuda.device_count()}):
    out = os.path.join(tmp, f"out{gpus}.tsv")
    t0 = time.time()
    p = subprocess.run([exe, "-l", lib, "-i", *paths, "-g", g2s, "-o", out, "-t", str(max(gpus, 2)), "--gpus", str(gpus),
                        "--timing"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert want in p.stderr, p.stderr
    t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
    tables[gpus] = open(out, "rb").read()
    print(t)
    print(f"gpus={gpus}: count_s {t['count_s']:.3f}  {t['reads'] / t['count_s'] / 1e6:.1f} M reads/s  wall {time.time() - t0:.1f} s  "
          f"rows {tables[gpus].count(10) - 1}", flush=True)
assert len(set(tables.values())) == 1, "tables differ between GPU counts"
print("tables identical; offsets as planted:", want)

# Config 5 in small: ONE sample (the first file grown to `one_reads` reads) on 1 GPU and cut into
# read shards over all visible GPUs (sgc_reduce_counts sums the shard vectors with NCCL): identical
# tables, the offset detected once.  Whole lines and span records both.
one_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * n_reads
big = os.path.join(tmp, "big.fastq.gz")
synth.Sample(seed, 100, arr, 75, 9, False).write_fastq(big, 0, one_reads, reads_per_member=1 << 20, gz_level=1)
single = {}
for gpus in sorted({1, torch.cuda.device_count()}):
    for extra in ([], ["--whole-lines"]):
        out = os.path.join(tmp, f"big{gpus}{len(extra)}.tsv")
        p = subprocess.run([exe, "-l", lib, "-i", big, "-o", out, "--gpus", str(gpus), "--read-shards", str(gpus), "--timing", *extra],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert p.stderr.count("Calculated Offsets: [Forward(9)]") == 1, p.stderr
        t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
        single[(gpus, len(extra))] = open(out, "rb").read()
        print(f"one sample, gpus={gpus} {'whole lines' if extra else 'span records'}: count_s {t['count_s']:.3f}  "
              f"{t['reads'] / t['count_s'] / 1e6:.1f} M reads/s  shards {t['read_shards_per_sample']}  "
              f"wait_inflate {t['wait_inflate_s']:.3f} copy {t['copy_to_pinned_s']:.3f} submit {t['submit_sync_s']:.3f} "
              f"tables {t['device_tables_s']:.2f} s", flush=True)
assert len(set(single.values())) == 1, "single-sample tables differ between GPU counts / framings"
print("single sample: tables identical on 1 and", torch.cuda.device_count(), "GPUs, offset detected once")

# The same 8 samples as BGZF: inflate and record framing run on the device that counts the sample,
# so the host's cores are out of the loop and the samples scale with the devices.
bpaths = []
for s, (rev, off) in enumerate(truth):
    p = os.path.join(tmp, f"b{s}.fastq.gz")
    synth.Sample(seed, s, arr, 75, off, rev).write_fastq_bgzf(p, 0, n_reads, gz_level=1)
    bpaths.append(p)
btables = {}
for gpus in sorted({1, torch.cuda.device_count()}):
    out = os.path.join(tmp, f"bout{gpus}.tsv")
    p = subprocess.run([exe, "-l", lib, "-i", *bpaths, "-g", g2s, "-o", out, "-t", str(max(gpus, 2)), "--gpus", str(gpus),
                        "--timing"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert want in p.stderr, p.stderr
    t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
    btables[gpus] = open(out, "rb").read()
    print(f"BGZF samples, gpus={gpus}: count_s {t['count_s']:.3f}  {t['reads'] / t['count_s'] / 1e6:.1f} M reads/s  "
          f"device_ingest_samples {t['device_ingest_samples']}  blocks {t['device_blocks']}", flush=True)
assert len(set(btables.values())) == 1
assert list(btables.values())[0].replace(b"\tb", b"\ts") == list(tables.values())[0], "BGZF and gzip tables differ"
print("BGZF samples: tables identical on 1 and", torch.cuda.device_count(), "GPUs and equal to the gzip samples' table")

# Config 5 end to end from gzip, on the device: ONE BGZF sample whose blocks the host cuts into waves at
# record boundaries; the waves go round the devices (inflate, framing and counting on each), the count
# vectors are summed at the end (peer copies; NCCL for inputs of 8 GiB or more).
bgzf_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 16 * n_reads
bigb = os.path.join(tmp, "bigb.fastq.gz")
t0 = time.time()
synth.Sample(seed, 100, arr, 75, 9, False).write_fastq_bgzf(bigb, 0, bgzf_reads, gz_level=1)
print(f"one BGZF sample of {bgzf_reads} reads: {os.path.getsize(bigb) / 1e6:.0f} MB written in {time.time() - t0:.1f} s", flush=True)
one_b = {}
for gpus in sorted({1, torch.cuda.device_count()}):
    out = os.path.join(tmp, f"bigb{gpus}.tsv")
    best = None
    for _ in range(2):
        p = subprocess.run([exe, "-l", lib, "-i", bigb, "-o", out, "--gpus", str(gpus), "--read-shards", str(gpus), "--timing"],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
        best = t if best is None or t["count_s"] < best["count_s"] else best
    assert p.stderr.count("Calculated Offsets: [Forward(9)]") == 1, p.stderr
    one_b[gpus] = open(out, "rb").read()
    print(f"one BGZF sample, gpus={gpus}: count_s {best['count_s']:.3f}  {best['reads'] / best['count_s'] / 1e6:.1f} M reads/s  "
          f"device_ingest_samples {best['device_ingest_samples']}  shards {best['read_shards_per_sample']}  blocks {best['device_blocks']}",
          flush=True)
assert len(set(one_b.values())) == 1, "tables differ between GPU counts"
print("one BGZF sample: tables identical on 1 and", torch.cuda.device_count(), "GPUs")
