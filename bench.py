#!/usr/bin/env python3
"""bench.py — reads/sec of the sgcount match-and-count path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Brunello-shaped synthetic library (77 441 x 20 bp, seed
0xB2000002), one sample of 50 M x 75 bp reads, one-mismatch table on, Forward(5).  A step is
one pass of the hot path over the whole sample.  With N > 1 (torchrun) every rank holds its
own 50 M-read shard of an N x 50 M-read sample (weak scaling) and the per-guide count vectors
are summed with an NCCL all-reduce inside the step.

Printed JSON (rank 0, one line):
  value      kernel-only reads/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through sgc_counter_submit with PINNED HOST buffers: H2D copies of the
             sequence lines and the D2H read-back of the count vector are inside the timed region
  roofline   algorithmic bytes (read_len+1 per read) / mean duration of the count launches
             against the measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's loop on this box's host cores (N = 1 only)
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0xB2000002
N_GUIDES = 77441
K = 20
READ_LEN = 75
OFFSET = 5
READS_PER_GPU = 50_000_000
WORKLOAD = "config2: Brunello-shaped 77441x20bp library, 1 sample x 50M x 75bp reads, 1-mismatch, Forward(5)"
FALLBACK_HBM_GBS = 6650.0


def cpu_threads() -> int:
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 64))
    except AttributeError:
        return max(1, min(os.cpu_count() or 1, 64))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# oracle leg: cpu_baseline of the GPU arm and the whole `--impl reference` arm
# ---------------------------------------------------------------------------------------------
class OracleLeg:
    """The reference is Rust and cannot be built in this image, so the CPU comparator is the
    oracle port (oracle/oracle.cpp): byte-string hash maps, literal Permuter, one token per
    probe.  The reference parallelises over samples only (count.rs:117-136); with one sample
    it is single-threaded.  We give it every host thread by sharding the sample's reads, which
    the reference itself cannot do — the number is generous to the CPU side."""

    def __init__(self, lib_arr, sample_reads: int, threads: int):
        from oracle import oracle as orc
        from sgcount_b200 import synth

        self.orc = orc
        self.threads = threads
        self.n = sample_reads
        recs = orc.Records.from_bytes(b"".join(b">lib.%d\n%s\n" % (i, lib_arr[i].tobytes()) for i in range(len(lib_arr))))
        self.lib_records = recs
        self.library = orc.Library.from_reader(recs)
        t0 = time.perf_counter()
        self.permuter = orc.Permuter.new(self.library)
        self.permuter_s = time.perf_counter() - t0
        sample = synth.Sample(SEED, 0, lib_arr, READ_LEN, OFFSET, False)
        self.lines = sample.fill_host(0, sample_reads)
        off = np.arange(0, self.lines.nbytes + 1, READ_LEN + 1, dtype=np.uint64)
        self.records = orc.Records.from_lines(self.lines, off)

    def step(self):
        t0 = time.perf_counter()
        c = self.orc.Counter.new(self.records, self.library, self.permuter, self.orc.Offset.Forward(OFFSET),
                                 None, True, n_threads=self.threads)
        dt = time.perf_counter() - t0
        return dt, c

    def describe(self):
        return (f"first {self.n} reads of the workload's sample 0, pre-parsed in memory, read-sharded over "
                f"{self.threads} threads; one-off Permuter::new took {self.permuter_s:.1f} s on 1 thread (not timed)")


# ---------------------------------------------------------------------------------------------
# end to end from a gzip FASTQ through the C++ host (ingest + count + table), with the oracle's
# gunzip + parse + match on the same file beside it
# ---------------------------------------------------------------------------------------------
def fastq_leg(lib_arr, n_reads: int, with_oracle: bool):
    import shutil
    import subprocess
    import tempfile

    from sgcount_b200 import synth

    exe = os.path.join(ROOT, "sgcount_b200", "lib", "sgcount")
    if not os.path.exists(exe):
        return {"unavailable": "sgcount host binary not built"}
    tmp = tempfile.mkdtemp(prefix="sgc_bench_")
    try:
        lib_path = os.path.join(tmp, "library.fa")
        with open(lib_path, "wb") as f:
            f.write(b"".join(b">lib.%d\n%s\n" % (i, lib_arr[i].tobytes()) for i in range(len(lib_arr))))
        fq = os.path.join(tmp, "sample0.fastq.gz")
        sample = synth.Sample(SEED, 0, lib_arr, READ_LEN, OFFSET, False)
        sample.write_fastq(fq, 0, n_reads, reads_per_member=1 << 20, gz_level=1)
        gz_bytes = os.path.getsize(fq)
        out_path = os.path.join(tmp, "counts.tsv")
        best = None
        for _ in range(2):  # the first run pages the file in
            t0 = time.perf_counter()
            p = subprocess.run([exe, "-l", lib_path, "-i", fq, "-a", str(OFFSET), "-q", "-o", out_path, "--timing"],
                               capture_output=True, text=True, timeout=900)
            wall = time.perf_counter() - t0
            if p.returncode != 0:
                return {"unavailable": "sgcount failed: " + p.stderr.strip()[-200:]}
            timing = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
            if best is None or timing["count_s"] < best[0]["count_s"]:
                best = (timing, wall)
        timing, wall = best
        rows = open(out_path).read().rstrip("\n").split("\n")
        table = {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows[1:]}
        res = {"value": timing["reads"] / timing["count_s"], "unit": "reads/s", "reads": timing["reads"],
               "count_s": timing["count_s"], "process_wall_s": wall, "gz_bytes": gz_bytes,
               "ingest_threads": timing["ingest_threads"],
               "members": (n_reads + (1 << 20) - 1) >> 20,
               "what": "sgcount CLI on a multi-member gzip FASTQ (1 Mi reads per member, so at most `members` inflate "
                       "threads have work): member-parallel inflate + record framing, H2D, count kernel, D2H; table "
                       "build and process start-up are outside count_s"}
        if with_oracle:
            from oracle import oracle as orc

            t0 = time.perf_counter()
            recs = orc.Records.from_path(fq)
            t_parse = time.perf_counter() - t0
            lib_recs = orc.Records.from_path(lib_path)
            olib = orc.Library.from_reader(lib_recs)
            operm = orc.Permuter.new(olib)
            t0 = time.perf_counter()
            oc = orc.Counter.new(recs, olib, operm, orc.Offset.Forward(OFFSET), None, True, n_threads=1)
            t_match = time.perf_counter() - t0
            res["cpu_port"] = {"value": n_reads / (t_parse + t_match), "unit": "reads/s", "cores": 1,
                               "gunzip_parse_s": t_parse, "match_s": t_match,
                               "note": "oracle port, one thread per sample like the reference (count.rs:117-136)"}
            counts = oc.counts_by_index()
            want = {"lib.%d" % i: int(c) for i, c in enumerate(counts) if c}
            res["parity"] = "ok" if want == table else "MISMATCH"
            assert want == table, "CLI count table differs from the oracle's"
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sgcount_b200 import synth

    threads = cpu_threads()
    lib_arr = synth.make_library(SEED, N_GUIDES, K)
    sample_reads = min(4_000_000, 250_000 * threads)
    leg = OracleLeg(lib_arr, sample_reads, threads)
    for _ in range(args.warmup):
        leg.step()
    times = [leg.step()[0] for _ in range(args.steps)]
    total = sum(times)
    value = sample_reads * args.steps / total
    out = {
        "impl": "reference",
        "metric": "reads/sec matched per B200 (kernel & end-to-end)",
        "value": value,
        "unit": "reads/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "reads_per_step": sample_reads, "read_len": READ_LEN, "n_guides": N_GUIDES},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": leg.describe()},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reason_bits = 0
        self.max_mhz = None
        self.active = threading.Event()
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        try:  # the first queries of a process are slow: take them before anything is timed
            for _ in range(3):
                self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            pass
        while not self.stop_flag:
            if self.active.is_set():
                try:
                    mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                    try:
                        bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append(mhz)
                    self.reason_bits |= int(bits)
                except Exception:
                    pass
                time.sleep(0.0005)  # an NVML query takes a fraction of a millisecond: sample back to back
            else:
                time.sleep(0.002)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        reasons = [name for bit, name in self.REASONS.items() if self.reason_bits & bit]
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(self.samples)}


BAD_REASONS = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import sgcount_b200 as sg
    from sgcount_b200 import _cabi, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Everything timed runs on ONE explicit (non-default) stream: the counter's kernels are
    # launched on it through the C ABI and the torch events / NCCL collectives are recorded on
    # it, so the CUDA events bracket exactly the launches they are meant to time.
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_reads = args.reads_per_gpu
    stride = READ_LEN + 1
    n_bytes = n_reads * stride

    # library + unified one-mismatch table on this GPU (replicated on every rank)
    lib_arr = synth.make_library(SEED, N_GUIDES, K)
    guides = [lib_arr[i].tobytes() for i in range(N_GUIDES)]
    library = sg.Library(guides, [b"lib.%d" % i for i in range(N_GUIDES)], device=local)
    permuter = sg.Permuter.new(library)
    info = permuter.info()

    # this rank's shard of sample 0, generated in HBM
    sample = synth.Sample(SEED, 0, lib_arr, READ_LEN, OFFSET, False)
    d_lines = torch.empty(n_bytes + 256, dtype=torch.uint8, device=dev)
    first = rank * n_reads
    sample.fill_device(first, n_reads, d_lines.data_ptr(), device=local, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()

    stream = work_stream.cuda_stream
    assert stream != 0 and torch.cuda.current_stream().cuda_stream == stream
    # Two counters on two state vectors: with N > 1 the all-reduce of step i runs on a second
    # stream while the kernel of step i+1 counts into the other vector, the way consecutive
    # samples of a real run overlap (sample i's reduce under sample i+1's counting).  Every
    # step's kernel AND reduce complete inside the timed region.
    n_buf = 2 if world > 1 else 1
    states = [torch.zeros(N_GUIDES + 2, dtype=torch.int64, device=dev) for _ in range(n_buf)]
    counters = [sg.Counter(library, permuter, sg.Offset.Forward(OFFSET), True, _cabi.RC_BITTRICK, stream=stream,
                           d_state=st.data_ptr()) for st in states]
    state, counter = states[0], counters[0]
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    reduce_done = [None] * n_buf

    def kernel_step(i, k_events=None):
        b = i % n_buf
        c, st = counters[b], states[b]
        if reduce_done[b] is not None:
            work_stream.wait_event(reduce_done[b])  # the vector's previous all-reduce has finished
        c.reset()
        if k_events is not None:
            k_events[0].record()
        c.submit_device(d_lines.data_ptr(), n_bytes, n_reads, stride, READ_LEN)
        if k_events is not None:
            k_events[1].record()
        if world > 1:
            counted = torch.cuda.Event()
            counted.record(work_stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(counted)
                shard.reduce_counts(st)  # NCCL all-reduce: the only state that crosses GPUs
                reduce_done[b] = torch.cuda.Event()
                reduce_done[b].record(comm_stream)

    def join_reduces():
        for ev in reduce_done:
            if ev is not None:
                work_stream.wait_event(ev)

    sampler = ClockSampler(local)
    sampler.start()

    def timed_kernel_arm():
        for i in range(args.warmup):
            kernel_step(i)
        join_reduces()
        barrier()
        launches0 = sum(c.launch_info().launches_total for c in counters)
        k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active.set()
        start.record()
        for i in range(args.steps):
            kernel_step(i, k_events[i])
        join_reduces()
        end.record()
        barrier()
        sampler.active.clear()
        total_ms = max_over_ranks(start.elapsed_time(end))
        kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in k_events)
        launches = sum(c.launch_info().launches_total for c in counters) - launches0
        return total_ms, kernel_ms, launches

    total_ms, kernel_ms, launches = timed_kernel_arm()
    if set(sampler.summary()["reasons"]) & BAD_REASONS:  # rejected: measure once more
        sampler.samples.clear()
        sampler.reason_bits = 0
        total_ms, kernel_ms, launches = timed_kernel_arm()
    clocks_kernel = sampler.summary()
    clocks_kernel["window"] = "timed region"
    if clocks_kernel["samples"] < 5:
        # K steps of ~1 ms are over before NVML answers a handful of queries: sample the SAME launches,
        # back to back for a quarter of a second, right after the timed region (not part of `value`)
        sampler.active.set()
        t_end = time.perf_counter() + 0.25
        i = 0
        while time.perf_counter() < t_end:
            for _ in range(20):
                kernel_step(i)
                i += 1
            join_reduces()
            torch.cuda.synchronize()
        barrier()
        sampler.active.clear()
        clocks_kernel = sampler.summary()
        clocks_kernel["window"] = "timed region + 0.25 s of the same launches right after it"
        kernel_step(args.steps - 1)  # leave the last timed step's table in its state vector
        join_reduces()
        torch.cuda.synchronize()

    last = (args.steps - 1) % n_buf
    state, counter = states[last], counters[last]
    counts_last, total_last, matched_last = counter.finish()
    value = world * n_reads * args.steps / (total_ms * 1e-3)

    # ---- end to end: pinned host lines -> sgc_counter_submit -> counts on the host -------------
    lib = _cabi.load()
    host_ptr = ctypes.c_void_p()
    _cabi.check(lib.sgc_host_alloc(ctypes.byref(host_ptr), n_bytes))
    sample.fill_host_ptr(first, n_reads, host_ptr.value)
    host_lines = np.ctypeslib.as_array(ctypes.cast(host_ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n_bytes,))
    batch = sg.ReadBatch(host_lines, n_reads, None, stride, READ_LEN)
    e2e_counter = sg.Counter(library, permuter, sg.Offset.Forward(OFFSET), True, _cabi.RC_BITTRICK, stream=stream,
                             d_state=state.data_ptr())
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        e2e_counter.reset()
        e2e_counter.submit(batch)
        if world > 1:
            shard.reduce_counts(state)
        return e2e_counter.finish()

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    sampler.samples.clear()
    sampler.active.set()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_result = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.active.clear()
    clocks_e2e = sampler.summary()
    sampler.stop_flag = True
    e2e_value = world * n_reads * e2e_steps / e2e_s
    del e2e_counter
    _cabi.check(lib.sgc_host_free(host_ptr))

    # the two arms must have produced the same table
    same = bool(np.array_equal(e2e_result[0], counts_last)) and e2e_result[1:] == (total_last, matched_last)
    assert total_last == world * n_reads, (total_last, world * n_reads)
    assert same, "kernel-only and end-to-end arms disagree"

    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        algo_bytes = n_reads * stride
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("count_stream_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": "reads/sec matched per B200 (kernel & end-to-end)",
            "value": value,
            "unit": "reads/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "reads_per_gpu": n_reads,
                "read_len": READ_LEN,
                "n_guides": N_GUIDES,
                "table_bytes": int(info.table_bytes),
                "n_variants": int(info.n_variants),
                "n_ambiguous": int(info.n_ambiguous),
                "table_build_ms": float(info.build_ms),
                "l2": "inputs (3.8 GB per step) are larger than L2; no flush needed",
                "parallelism": (f"read-sharded x{world}, NCCL all-reduce of u64[{N_GUIDES + 2}] per step on a second stream, "
                                "overlapping the next step's kernel") if world > 1 else "1 GPU",
                "matched_fraction": matched_last / max(total_last, 1),
            },
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": n_bytes,
                    "d2h_bytes_per_step": (N_GUIDES + 2) * 8, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "note": "kernel = count_stream_kernel (the only count launch of a step: 50 M reads are whole 32-read tiles); "
                                 "duration = CUDA events around sgc_counter_submit_device on the launching stream, mean over the timed steps"},
            "clocks": {"sm_mhz": clocks_kernel["sm_mhz"], "sm_max_mhz": clocks_kernel["sm_max_mhz"],
                       "reasons": clocks_kernel["reasons"], "samples": clocks_kernel["samples"],
                       "window": clocks_kernel["window"], "e2e_sm_mhz": clocks_e2e["sm_mhz"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = cpu_threads()
            sample_reads = min(4_000_000, 250_000 * threads, n_reads)
            leg = OracleLeg(lib_arr, sample_reads, threads)
            dt, oc = leg.step()
            out["cpu_baseline"] = {"value": sample_reads / dt, "unit": "reads/s", "cores": threads, "kind": "port",
                                   "sample": leg.describe()}
            # parity of the bench's own run: the GPU on the same sample vs the oracle
            chk = sg.Counter(library, permuter, sg.Offset.Forward(OFFSET), True, _cabi.RC_BITTRICK, stream=stream)
            chk.submit_device(d_lines.data_ptr(), sample_reads * stride, sample_reads, stride, READ_LEN)
            g_counts, g_total, g_matched = chk.finish()
            ok = (np.array_equal(g_counts, oc.counts_by_index()) and g_total == oc.total_reads()
                  and g_matched == oc.matched_reads())
            out["parity"] = "ok" if ok else "MISMATCH"
            assert ok, "GPU counts differ from the oracle on the cpu_baseline sample"
        if world == 1 and args.fastq_reads > 0:
            out["e2e_fastq"] = fastq_leg(lib_arr, args.fastq_reads, with_oracle=not args.no_cpu_baseline)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=READS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fastq-reads", type=int, default=16 << 20,
                    help="reads of the gzip-FASTQ end-to-end leg through the C++ host (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
