"""BASELINE configs 3-5 at their library sizes and reduced read counts, end to end through the
C++ host: synthetic multi-member gzip FASTQ -> auto-detected offsets (forward and reverse) ->
one-mismatch counting -> count table, compared with the oracle's table on the same files.
(Config 1 is tests/test_gpu_parity.py + test_host_cli.py, config 2 test_gpu_parity.py's
Brunello-shaped run and bench.py's own parity check.)"""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from sgcount_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sgcount_b200", "lib", "sgcount")


def write_library(path, arr, with_genes=None):
    with open(path, "wb") as f:
        f.write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
    if with_genes:
        with open(with_genes, "wb") as f:
            f.write(b"".join(b"gene.%d\tlib.%d\n" % (i // 10, i) for i in range(len(arr))))


def run_cli(*args):
    p = subprocess.run([BIN, *args], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr
    return p


def oracle_table(lib_path, paths, names, offsets, g2s=None):
    lib_recs = orc.Records.from_path(lib_path)
    olib = orc.Library.from_reader(lib_recs)
    operm = orc.Permuter.new(olib)
    counters = []
    for path, off in zip(paths, offsets):
        recs = orc.Records.from_path(path)
        detected = orc.entropy_offset(lib_recs, recs, 5000)
        assert (detected.reverse, detected.index) == off, (path, detected)
        counters.append(orc.Counter.new(recs, olib, operm, detected, None, True, n_threads=os.cpu_count() or 4))
    text = orc.render_results(counters, names, olib, open(g2s, "rb").read() if g2s else None, include_zero=False)
    return counters, text


def table(text):
    lines = text.rstrip("\n").split("\n")
    return lines[0], sorted(lines[1:])


def test_config3_four_samples_mixed_offsets(tmp_path):
    """GeCKO-v2-shaped library (123 411 guides), truth F(0), R(12), F(23), R(5); 100 k reads each"""
    seed = 0xB2000003
    arr = synth.make_library(seed, 123411, 20)
    lib_path = str(tmp_path / "gecko.fa")
    write_library(lib_path, arr)
    truth = [(False, 0), (True, 12), (False, 23), (True, 5)]
    paths, names = [], []
    for s, (rev, off) in enumerate(truth):
        p = str(tmp_path / f"s{s}.fastq.gz")
        synth.Sample(seed, s, arr, 75, off, rev).write_fastq(p, 0, 100_000, reads_per_member=30_000)
        paths.append(p)
        names.append(f"s{s}")
    cli = run_cli("-l", lib_path, "-i", *paths, "-t", "4", "-o", str(tmp_path / "out.tsv"))
    assert "Calculated Offsets: [Forward(0), Reverse(12), Forward(23), Reverse(5)]" in cli.stderr
    counters, want = oracle_table(lib_path, paths, names, truth)
    assert table(open(tmp_path / "out.tsv").read()) == table(want)
    for name, c in zip(names, counters):
        assert f"Finished: {name}; Fraction mapped: {c.matched_reads() / c.total_reads():.3f} " \
               f"[{c.matched_reads()} / {c.total_reads()}]" in cli.stderr


def test_config4_gene_map_and_config5_read_sharding(tmp_path):
    """CRISPRi-shaped library (200 000 guides, 20 000 genes) with a gene map; one sample is also
    counted as four read shards through the C ABI and summed (config 5's partitioning)."""
    import torch

    import sgcount_b200 as sg
    from sgcount_b200 import shard

    seed = 0xB2000004
    arr = synth.make_library(seed, 200000, 20)
    lib_path, g2s = str(tmp_path / "crispri.fa"), str(tmp_path / "g2s.txt")
    write_library(lib_path, arr, g2s)
    truth = [(False, 7), (True, 30)]
    paths, names = [], ["a", "b"]
    for s, (rev, off) in enumerate(truth):
        p = str(tmp_path / f"{names[s]}.fq.gz")
        synth.Sample(seed, s, arr, 75, off, rev).write_fastq(p, 0, 60_000, reads_per_member=25_000)
        paths.append(p)
    run_cli("-l", lib_path, "-i", *paths, "-g", g2s, "-o", str(tmp_path / "out.tsv"), "-q")
    counters, want = oracle_table(lib_path, paths, names, truth, g2s)
    got = open(tmp_path / "out.tsv").read()
    assert got.split("\n")[0] == "Guide\tGene\ta\tb"
    assert table(got) == table(want)

    # config 5: the same reads cut into 4 shards, one counter each, state vectors summed
    library = sg.Library([arr[i].tobytes() for i in range(len(arr))], [b"lib.%d" % i for i in range(len(arr))])
    permuter = sg.Permuter.new(library)
    sample = synth.Sample(seed, 0, arr, 75, 7, False)
    n = 60_000
    lines = sample.fill_host(0, n)
    off = sg.entropy_offset(library, sg.ReadBatch(lines, n, None, 76, 75), 5000)
    assert off == sg.Offset.Forward(7)
    total = torch.zeros(len(arr) + 2, dtype=torch.int64, device="cuda")
    for sh in shard.plan_shards([n], 4):
        state = torch.zeros(len(arr) + 2, dtype=torch.int64, device="cuda")
        c = sg.Counter(library, permuter, off, True, d_state=state.data_ptr())
        part = sg.ReadBatch(lines[sh.first_read * 76:(sh.first_read + sh.n_reads) * 76], sh.n_reads, None, 76, 75)
        c.submit(part)
        c.sync()
        total += shard.reduce_counts(state)
        del c
    total = total.cpu().numpy()
    assert np.array_equal(total[:-2].astype(np.uint64), counters[0].counts_by_index())
    assert (int(total[-2]), int(total[-1])) == (counters[0].total_reads(), counters[0].matched_reads())


def test_config2_full_size_properties():
    """Config 2 at its FULL size (77 441 guides, 50 M x 75 bp reads generated in HBM): the oracle
    cannot finish 50 M reads in a test, so the whole-sample table is pinned through
    size-independent properties: totals, additivity over read shards (each shard small enough to
    be oracle-checked), and a second pass doubling every count."""
    import torch

    import sgcount_b200 as sg

    seed, n_guides, n_reads, stride = 0xB2000002, 77441, 50_000_000, 76
    arr = synth.make_library(seed, n_guides, 20)
    library = sg.Library([arr[i].tobytes() for i in range(n_guides)], [b"lib.%d" % i for i in range(n_guides)])
    permuter = sg.Permuter.new(library)
    sample = synth.Sample(seed, 0, arr, 75, 5, False)
    d = torch.empty(n_reads * stride + 256, dtype=torch.uint8, device="cuda")
    sample.fill_device(0, n_reads, d.data_ptr())
    torch.cuda.synchronize()

    whole = sg.Counter(library, permuter, sg.Offset.Forward(5))
    whole.submit_device(d.data_ptr(), n_reads * stride, n_reads, stride, 75)
    counts, total, matched = whole.finish()
    assert total == n_reads and matched == int(counts.sum()) and 0.9 < matched / total < 0.97
    assert whole.launch_info().kernel == 0  # the streaming kernel

    # additivity: 25 shards of 2 M reads (ragged last tile boundaries included) sum to the whole;
    # the first shard is checked against the oracle read by read
    acc = np.zeros(n_guides, dtype=np.uint64)
    acc_total = acc_matched = 0
    bounds = [0] + [2_000_000 * i + 17 * i for i in range(1, 25)] + [n_reads]
    for a, b in zip(bounds, bounds[1:]):
        c = sg.Counter(library, permuter, sg.Offset.Forward(5))
        c.submit_device(d.data_ptr() + a * stride, (b - a) * stride, b - a, stride, 75)
        sc, st, sm = c.finish()
        acc += sc.astype(np.uint64)
        acc_total += st
        acc_matched += sm
        if a == 0:
            lib_recs = orc.Records.from_bytes(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(n_guides)))
            olib = orc.Library.from_reader(lib_recs)
            lines = sample.fill_host(0, b)
            recs = orc.Records.from_lines(lines, np.arange(0, lines.nbytes + 1, stride, dtype=np.uint64))
            oc = orc.Counter.new(recs, olib, orc.Permuter.new(olib), orc.Offset.Forward(5), None, True,
                                 n_threads=os.cpu_count() or 4)
            assert np.array_equal(sc.astype(np.uint64), oc.counts_by_index())
            assert (st, sm) == (oc.total_reads(), oc.matched_reads())
        del c
    assert np.array_equal(acc, counts.astype(np.uint64)) and (acc_total, acc_matched) == (total, matched)

    # a second pass over the same reads doubles every counter (no state leaks between launches)
    whole.submit_device(d.data_ptr(), n_reads * stride, n_reads, stride, 75)
    counts2, total2, matched2 = whole.finish()
    assert np.array_equal(counts2, 2 * counts) and (total2, matched2) == (2 * total, 2 * matched)

    # ... and so do 18 more: a hand-back of a ring buffer that races with the lanes still reading
    # it loses a read in a hundred million, which only repetition at full size shows
    for _ in range(18):
        whole.submit_device(d.data_ptr(), n_reads * stride, n_reads, stride, 75)
    counts20, total20, matched20 = whole.finish()
    assert np.array_equal(counts20, 20 * counts) and (total20, matched20) == (20 * total, 20 * matched)
