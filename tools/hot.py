#!/usr/bin/env python3
"""Kernel time when a fraction of the reads all carry the SAME guide (positive-selection screens):
every one of them is an atomic on one counter.  Tuning aid."""
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
d = torch.empty(N * 76 + 256, dtype=torch.uint8, device="cuda")
for frac in [float(x) for x in (sys.argv[1:] or ["0", "0.01", "0.1", "0.3", "0.9"])]:
    sample.fill_device(0, N, d.data_ptr())
    torch.cuda.synchronize()
    rows = d[:N * 76].view(N, 76)
    if frac > 0:
        hot = rows[0].clone()
        hot[5:25] = torch.from_numpy(arr[12345].copy()).cuda()
        mask = torch.rand(N, device="cuda") < frac
        rows[mask] = hot
        del mask
    ref = None
    for plan, replicas in (("auto", 0), ("off", 1), ("16 replicas", 16)):
        counter = sg.Counter(library, permuter, sg.Offset.Forward(5))
        counter.set_replicas(replicas)
        t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0[0].record()
        counter.submit_device(d.data_ptr(), N * 76, N, 76, 75)  # the first launch carries the plan
        t0[1].record()
        counter.submit_device(d.data_ptr(), N * 76, N, 76, 75)
        torch.cuda.synchronize()
        counter.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            counter.submit_device(d.data_ptr(), N * 76, N, 76, 75)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        counts, total, matched = counter.finish()
        li = counter.launch_info()
        ref = counts if ref is None else ref
        print(f"hot fraction {frac:.2f} plan={plan:<11} replicas={li.replicas} hot={li.hot_guides}: {ms:.3f} ms  "
              f"{N / ms / 1e6:.2f} Greads/s  frac={N * 76 / ms / 1e6 / 6547.2:.3f}  first launch {t0[0].elapsed_time(t0[1]):.3f} ms  "
              f"max count share={counts.max() / max(total, 1):.3f} same={bool((counts == ref).all())}", flush=True)
