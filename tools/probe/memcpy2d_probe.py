#!/usr/bin/env python3
"""Tuning aid: can the copy engine gather the guide-window span (28 of 76 bytes per line) from
pinned host memory faster than it copies whole lines?  cudaMemcpy2DAsync host->device with a source
pitch of 76 and a width of 28..76 bytes."""
import time

import torch
from cuda import cudart

n = 50_000_000
pitch = 76
err, host = cudart.cudaHostAlloc(n * pitch, 0)
assert err == cudart.cudaError_t.cudaSuccess, err
dev = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
err, stream = cudart.cudaStreamCreate()
for width, dpitch in ((76, 76), (32, 32), (28, 32), (24, 24), (64, 64)):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        (err,) = cudart.cudaMemcpy2DAsync(dev.data_ptr(), dpitch, host, pitch, width, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, stream)
        assert err == cudart.cudaError_t.cudaSuccess, err
        cudart.cudaStreamSynchronize(stream)
        best = min(best, time.perf_counter() - t0)
    print(f"width {width:2d} of pitch {pitch}: {best * 1e3:7.1f} ms  {n * width / best / 1e9:6.1f} GB/s of payload", flush=True)
