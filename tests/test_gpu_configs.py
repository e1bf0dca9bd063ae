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
