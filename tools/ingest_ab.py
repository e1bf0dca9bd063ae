#!/usr/bin/env python3
"""End-to-end gzip-FASTQ ingest of the `sgcount` CLI with the host's own gzip decoder and with
zlib (SGC_INFLATE=zlib), same file, same box.  Tuning aid."""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgcount_b200 import synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 32 << 20
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sgcount_b200", "lib", "sgcount")
arr = synth.make_library(0xB2000002, 77441, 20)
tmp = tempfile.mkdtemp(prefix="sgc_ingest_")
lib = os.path.join(tmp, "lib.fa")
open(lib, "wb").write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
fq = os.path.join(tmp, "s.fastq.gz")
t0 = time.time()
synth.Sample(0xB2000002, 0, arr, 75, 5, False).write_fastq(fq, 0, n_reads, reads_per_member=1 << 20, gz_level=1)
print(f"generated {n_reads} reads, {os.path.getsize(fq) / 1e6:.0f} MB gz in {time.time() - t0:.1f} s", flush=True)
for rep in range(2):
    for mode in ("fast", "zlib"):
        env = dict(os.environ)
        if mode == "zlib":
            env["SGC_INFLATE"] = "zlib"
        p = subprocess.run([exe, "-l", lib, "-i", fq, "-a", "5", "-q", "-o", os.path.join(tmp, "o.tsv"), "--timing"],
                           capture_output=True, text=True, env=env)
        t = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
        print(f"{mode}: count_s {t['count_s']:.3f}  {t['reads'] / t['count_s'] / 1e6:.1f} M reads/s  threads {t['ingest_threads']}  wait {t['wait_inflate_s']:.3f} copy {t['copy_to_pinned_s']:.3f} submit {t['submit_sync_s']:.3f}", flush=True)
