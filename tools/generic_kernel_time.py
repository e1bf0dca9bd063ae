#!/usr/bin/env python3
"""Kernel-only time of the GENERIC count kernel (variable-length lines through u32 offsets, or an
unaligned fixed-stride buffer) on the bench workload, next to the streaming kernel.  Tuning aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sgcount_b200 as sg
from sgcount_b200 import synth

N = int(os.environ.get("TUNE_READS", 50_000_000))
arr = synth.make_library(0xB2000002, 77441, 20)
library = sg.Library([arr[i].tobytes() for i in range(len(arr))], [b"g%d" % i for i in range(len(arr))])
permuter = sg.Permuter.new(library)
sample = synth.Sample(0xB2000002, 0, arr, 75, 5, False)
d = torch.empty(N * 76 + 512, dtype=torch.uint8, device="cuda")
sample.fill_device(0, N, d.data_ptr())
off = (torch.arange(N + 1, dtype=torch.int64, device="cuda") * 76).to(torch.int32)  # < 4 GiB
torch.cuda.synchronize()


def timed(label, fn, iters=5):
    c = sg.Counter(library, permuter, sg.Offset.Forward(5))
    for _ in range(2):
        fn(c)
    torch.cuda.synchronize()
    c.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn(c)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    counts, total, matched = c.finish()
    print(f"{label}: {ms:.3f} ms  {N / ms / 1e6:.2f} Greads/s  frac={N * 76 / ms / 1e6 / 6547.2:.3f}  matched/iter={matched // iters} "
          f"kernel={c.launch_info().kernel}", flush=True)
    return counts // iters


ref = timed("streaming (fixed stride, aligned)", lambda c: c.submit_device(d.data_ptr(), N * 76, N, 76, 75))
got = timed("generic (u32 line offsets)", lambda c: c.submit_device(d.data_ptr(), N * 76, N, 0, 0, off.data_ptr()))
assert (ref == got).all()
shifted = d[4:4 + N * 76].clone() if False else None
