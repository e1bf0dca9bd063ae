"""Oracle vs the matcher-independent labels of the example fixtures (config 1).

expected.json is a tally of the read headers (tests/golden/make_golden.py), not the output
of any matcher, so agreement here pins the oracle's whole count path on real data:
gzip + FASTQ framing, entropy offset detection, trimming, lookup, fold and the table.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as orc

FIXTURES = ["sequence", "zero.sequence", "diff.sequence", "offset", "offset_clipped"]


@pytest.fixture(scope="module")
def library_records(example_dir):
    return orc.Records.from_path(os.path.join(example_dir, "library.fasta.gz"))


@pytest.fixture(scope="module")
def library(library_records):
    return orc.Library.from_reader(library_records)


def test_library_shape(library, expected):
    assert len(library) == expected["library"]["n"] == 100
    assert library.size() == expected["library"]["k"] == 20
    assert [a.decode() for a in library.values()] == expected["library"]["aliases"]


@pytest.mark.parametrize("name", FIXTURES)
def test_entropy_offset_is_forward_5(name, example_dir, library_records):
    sample = orc.Records.from_path(os.path.join(example_dir, name + ".fastq.gz"))
    assert orc.entropy_offset(library_records, sample, 5000) == orc.Offset.Forward(5)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("name", FIXTURES)
def test_counts_equal_header_labels(name, exact, example_dir, library, expected):
    fx = expected["fixtures"][name]
    sample = orc.Records.from_path(os.path.join(example_dir, name + ".fastq.gz"))
    assert len(sample) == fx["total_reads"]
    permuter = None if exact else orc.Permuter.new(library)
    counter = orc.Counter.new(sample, library, permuter, orc.Offset.Forward(5), 20, True)
    assert counter.total_reads() == fx["total_reads"]
    assert counter.matched_reads() == fx["matched_reads"]
    assert counter.counts_by_index().tolist() == fx["counts"]
    rows = "\n".join(f"{a}\t{c}" for a, c in sorted(zip(expected["library"]["aliases"], fx["counts"])))
    assert hashlib.sha256(rows.encode()).hexdigest()[:16] == fx["sha256_16"]


def test_example_permuter_has_no_ambiguity(library):
    """min pairwise Hamming distance of the example library is 7 -> 80 variants per guide"""
    p = orc.Permuter.new(library)
    assert p.map_len() == 100 * 20 * 4
    assert p.null_len() == 100


def test_results_table(example_dir, library, expected):
    names = ["sequence", "zero.sequence"]
    counters = [orc.Counter.new(orc.Records.from_path(os.path.join(example_dir, n + ".fastq.gz")), library,
                                None, orc.Offset.Forward(5)) for n in names]
    g2s = open(os.path.join(example_dir, "g2s.txt"), "rb").read()
    text = orc.render_results(counters, names, library, g2s, include_zero=False)
    lines = text.rstrip("\n").split("\n")
    assert lines[0] == "Guide\tGene\tsequence\tzero.sequence"
    assert len(lines) == 101
    a = expected["fixtures"]["sequence"]["counts"]
    b = expected["fixtures"]["zero.sequence"]["counts"]
    assert lines[1] == f"lib.0\tgene.0\t{a[0]}\t{b[0]}"
    assert lines[100] == f"lib.99\tgene.9\t{a[99]}\t{b[99]}"
    # zero rows vanish unless -z (results.rs:90-94)
    only_zero = orc.render_results(counters[1:], names[1:], library, None, include_zero=False)
    assert len(only_zero.rstrip("\n").split("\n")) == 1 + 90
    with_zero = orc.render_results(counters[1:], names[1:], library, None, include_zero=True)
    assert len(with_zero.rstrip("\n").split("\n")) == 1 + 100
    assert with_zero.split("\n")[0] == "Guide\tzero.sequence"


def test_sample_names():
    """utils.rs:18-49 (+ its tests at utils.rs:60-104)"""
    assert orc.generate_sample_names(["a/b/x.fastq.gz", "y.fq", "z.fasta", "w.fa.gz"]) == ["x", "y", "z", "w"]
    assert orc.generate_sample_names(["a/x.fq", "b/x.fq.gz"]) == ["Sample.0", "Sample.1"]
