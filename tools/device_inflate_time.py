#!/usr/bin/env python3
"""Times the device ingest (sgc_fastq_stream_*: BGZF inflate + record framing + count) on a synthetic
BGZF FASTQ of the bench workload, next to the host path of the CLI on the same file.  Tuning aid."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import sgcount_b200 as sg
from sgcount_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
WAVE = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
arr = synth.make_library(0xB2000002, 77441, 20)
library = sg.Library([arr[i].tobytes() for i in range(len(arr))], [b"lib.%d" % i for i in range(len(arr))])
permuter = sg.Permuter.new(library)
sample = synth.Sample(0xB2000002, 0, arr, 75, 5, False)
tmp = tempfile.mkdtemp(prefix="sgc_dinf_")
path = os.path.join(tmp, "s.fastq.gz")
t0 = time.time()
sample.write_fastq_bgzf(path, 0, N)
print(f"wrote {N} reads as BGZF: {os.path.getsize(path) / 1e6:.0f} MB in {time.time() - t0:.1f} s", flush=True)
blob = np.fromfile(path, dtype=np.uint8)
t0 = time.time()
begin, isize = sg.bgzf_blocks(blob.tobytes())
print(f"{len(isize)} blocks, {int(isize.sum()) / 1e9:.2f} GB of text; block walk in Python {time.time() - t0:.1f} s", flush=True)
start, length, span_off = sg.span_geometry(20, 75, sg.Offset.Forward(5), True)
ref = None
for rep in range(3):
    c = sg.Counter(library, permuter, span_off, True)
    stream = sg.FastqStream(c, 75, start, length)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a in range(0, len(isize), WAVE):
        b = min(len(isize), a + WAVE)
        stream.submit(blob, begin[a:b + 1], isize[a:b])
    n = stream.finish()
    counts, total, matched = c.finish()
    dt = time.perf_counter() - t0
    print(f"device ingest: {n} records in {dt * 1e3:.1f} ms = {n / dt / 1e6:.1f} M reads/s  "
          f"({int(isize.sum()) / dt / 1e9:.1f} GB/s of text, {blob.nbytes / dt / 1e9:.2f} GB/s compressed)  matched {matched}", flush=True)
    ref = counts if ref is None else ref
    assert (counts == ref).all() and total == N
if os.environ.get("DINF_SKIP_CLI"):
    sys.exit(0)
lib_path = os.path.join(tmp, "lib.fa")
open(lib_path, "wb").write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
exe = os.path.join(ROOT, "sgcount_b200", "lib", "sgcount")
for extra in ([], ["--host-inflate"]):
    for _ in range(2):
        p = subprocess.run([exe, "-l", lib_path, "-i", path, "-a", "5", "-q", "-o", os.path.join(tmp, "o.tsv"), "--timing", *extra],
                           capture_output=True, text=True)
    print("CLI", extra, p.stderr.strip().splitlines()[-1] if p.returncode == 0 else p.stderr[-300:], flush=True)
