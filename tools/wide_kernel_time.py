#!/usr/bin/env python3
"""Kernel-only time of the streaming kernel at other guide lengths (k = 16 narrow/8-word, 24 and
30 wide) on Brunello-sized synthetic libraries.  Tuning aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sgcount_b200 as sg
from sgcount_b200 import synth

N = int(os.environ.get("TUNE_READS", 50_000_000))
d = torch.empty(N * 76 + 512, dtype=torch.uint8, device="cuda")
for k in [int(x) for x in os.environ.get('WIDE_KS', '16,20,24,28,30').split(',')]:
    arr = synth.make_library(0xB2000002 + k, 77441, k)
    library = sg.Library([arr[i].tobytes() for i in range(len(arr))], [b"g%d" % i for i in range(len(arr))])
    permuter = sg.Permuter.new(library)
    sample = synth.Sample(0xB2000002 + k, 0, arr, 75, 5, False)
    sample.fill_device(0, N, d.data_ptr())
    torch.cuda.synchronize()
    c = sg.Counter(library, permuter, sg.Offset.Forward(5))
    for _ in range(2):
        c.submit_device(d.data_ptr(), N * 76, N, 76, 75)
    torch.cuda.synchronize()
    c.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        c.submit_device(d.data_ptr(), N * 76, N, 76, 75)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    counts, total, matched = c.finish()
    info = permuter.info()
    print(f"k={k}: {ms:.3f} ms  {N / ms / 1e6:.2f} Greads/s  frac={N * 76 / ms / 1e6 / 6547.2:.3f}  matched={matched / total:.4f} "
          f"tables={info.table_bytes / 1e6:.1f}MB kernel={c.launch_info().kernel} smem={c.launch_info().smem_bytes}", flush=True)
