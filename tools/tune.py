#!/usr/bin/env python3
"""Times the count kernel on the bench workload for a list of ring configurations
(SGC_WARPS/SGC_CTAS/SGC_STAGES are read at every launch).  Tuning aid, not a benchmark."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the switches below exist only in the tuning build of the C ABI (make -C sgcount_b200/csrc tuning)
if not os.environ.get("SGC_CUDA_LIB"):
    import subprocess

    subprocess.run(["make", "-C", os.path.join(ROOT, "sgcount_b200", "csrc"), "tuning"], check=True, capture_output=True)
    os.environ["SGC_CUDA_LIB"] = os.path.join(ROOT, "sgcount_b200", "lib", "libsgcount_cuda_tuning.so")
import torch

import sgcount_b200 as sg
from sgcount_b200 import synth

N_READS = int(os.environ.get("TUNE_READS", 50_000_000))
lib_arr = synth.make_library(0xB2000002, 77441, 20)
library = sg.Library([lib_arr[i].tobytes() for i in range(len(lib_arr))], [b"g%d" % i for i in range(len(lib_arr))])
permuter = sg.Permuter.new(library)
sample = synth.Sample(0xB2000002, 0, lib_arr, 75, 5, False)
d = torch.empty(N_READS * 76 + 256, dtype=torch.uint8, device="cuda")
sample.fill_device(0, N_READS, d.data_ptr())
torch.cuda.synchronize()
counter = sg.Counter(library, permuter, sg.Offset.Forward(5))
configs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[1:]] or [(12, 2, 3)]
configs = [c + (0, -1)[len(c) - 3:] for c in configs]  # warps,ctas,stages[,debug[,carveout %]]
ref = None
for warps, ctas, stages, debug, carve in configs:
    os.environ.update(SGC_WARPS=str(warps), SGC_CTAS=str(ctas), SGC_STAGES=str(stages), SGC_DEBUG=str(debug),
                      SGC_CARVEOUT=str(carve))
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
    li = counter.launch_info()
    if ref is None and debug == 0:
        ref = counts.copy()
    ok = ref is not None and (counts == ref).all()
    if debug == 0 and N_READS == 50_000_000 and matched != 46992470 * iters:
        # the oracle-checked total of this workload: anything else is a lost or double-counted read
        print(f"!!! matched {matched} != {46992470 * iters}: WRONG RESULT", flush=True)
    print(f"warps={warps} ctas={ctas} stages={stages} debug={debug} carve={carve} grid={li.grid} smem={li.smem_bytes} "
          f"{ms:.3f} ms  {N_READS/ms/1e6:.2f} Greads/s  {N_READS*76/ms/1e6:.0f} GB/s "
          f"frac={N_READS*76/ms/1e6/6547.2:.3f} matched={matched//iters} same={ok}", flush=True)
