#!/usr/bin/env python3
"""Kernel-only time of the streaming kernel at other read lengths (50, 100, 150, 250 bp).  Tuning aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sgcount_b200 as sg
from sgcount_b200 import synth

arr = synth.make_library(0xB2000002, 77441, 20)
library = sg.Library([arr[i].tobytes() for i in range(len(arr))], [b"g%d" % i for i in range(len(arr))])
permuter = sg.Permuter.new(library)
for L in (50, 75, 100, 150, 250):
    N = int(3.8e9 // (L + 1)) // 32 * 32
    d = torch.empty(N * (L + 1) + 512, dtype=torch.uint8, device="cuda")
    sample = synth.Sample(0xB2000002, 0, arr, L, 5, False)
    sample.fill_device(0, N, d.data_ptr())
    torch.cuda.synchronize()
    c = sg.Counter(library, permuter, sg.Offset.Forward(5))
    for _ in range(2):
        c.submit_device(d.data_ptr(), N * (L + 1), N, L + 1, L)
    torch.cuda.synchronize()
    c.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        c.submit_device(d.data_ptr(), N * (L + 1), N, L + 1, L)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    counts, total, matched = c.finish()
    li = c.launch_info()
    print(f"L={L}: {N} reads {ms:.3f} ms  {N / ms / 1e6:.2f} Greads/s  frac={N * (L + 1) / ms / 1e6 / 6547.2:.3f}  matched={matched / total:.4f} "
          f"grid={li.grid} block={li.block} smem={li.smem_bytes}", flush=True)
    del d, c
