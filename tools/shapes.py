#!/usr/bin/env python3
"""Kernel-only timing of the count path on the other BASELINE shapes (library size, orientation,
offset), checked against each other where the answer must agree.  Tuning aid, not the benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sgcount_b200 as sg
from sgcount_b200 import synth

N_READS = int(os.environ.get("TUNE_READS", 50_000_000))
CASES = [  # seed, n_guides, reverse, offset, with_perm, recursion
    (0xB2000002, 77441, False, 5, True, True),
    (0xB2000002, 77441, True, 12, True, True),
    (0xB2000002, 77441, False, 5, False, True),
    (0xB2000002, 77441, False, 5, True, False),
    (0xB2000003, 123411, True, 5, True, True),
    (0xB2000003, 123411, False, 23, True, True),
    (0xB2000004, 200000, False, 0, True, True),
    (0xB2000004, 200000, True, 17, True, True),
]
d = torch.empty(N_READS * 76 + 256, dtype=torch.uint8, device="cuda")
libs = {}
for seed, n, rev, off, perm, rec in CASES:
    if (seed, n) not in libs:
        arr = synth.make_library(seed, n, 20)
        library = sg.Library([arr[i].tobytes() for i in range(n)], [b"g%d" % i for i in range(n)])
        libs[(seed, n)] = (arr, library, sg.Permuter.new(library))
    arr, library, permuter = libs[(seed, n)]
    sample = synth.Sample(seed, 1, arr, 75, off, rev)
    sample.fill_device(0, N_READS, d.data_ptr())
    torch.cuda.synchronize()
    counter = sg.Counter(library, permuter if perm else None, sg.Offset(rev, off), rec)
    for _ in range(3):
        counter.submit_device(d.data_ptr(), N_READS * 76, N_READS, 76, 75)
    torch.cuda.synchronize()
    counter.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    a.record()
    for _ in range(iters):
        counter.submit_device(d.data_ptr(), N_READS * 76, N_READS, 76, 75)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    counts, total, matched = counter.finish()
    info = permuter.info()
    print(f"guides={n} {'Reverse' if rev else 'Forward'}({off}) perm={int(perm)} recursion={int(rec)} "
          f"{ms:.3f} ms {N_READS / ms / 1e6:.2f} Greads/s frac={N_READS * 76 / ms / 1e6 / 6547.2:.3f} "
          f"matched={matched / total:.4f} tables={info.table_bytes / 1e6:.1f}MB left_out={info.front_left_out}", flush=True)
