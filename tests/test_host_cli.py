"""The C++ host `sgcount` (sgcount_b200/host): the reference's command line over the C ABI.

CPU part: flag surface and the validations the reference performs before it touches a read
(main.rs:130-160, count.rs:62-100, library.rs:83, genemap.rs:53-68).  GPU part: whole runs
on the example fixtures compared with the oracle's table (results.rs:71-99 restated in
oracle.render_results) and the golden counts."""
import gzip
import os
import subprocess

import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sgcount_b200", "lib", "sgcount")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(BIN):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(BIN)


def run(*args, check=False):
    p = subprocess.run([BIN, *args], capture_output=True, text=True, timeout=300)
    if check:
        assert p.returncode == 0, p.stderr
    return p


def test_help_lists_the_reference_flags():
    p = run("--help", check=True)
    for flag in ["-l, --library-path", "-i, --input-paths", "-n, --sample-names", "-o, --output-path", "-g, --genemap",
                 "-a, --offset", "-p, --no-position-recursion", "-r, --reverse", "-x, --exact", "-s, --subsample",
                 "-t, --threads", "-q, --quiet", "-z, --include-zero"]:
        assert flag in p.stdout


def test_required_arguments_and_unknown_flags():
    assert run().returncode == 1
    assert "--library-path" in run("-i", "x").stderr
    assert "--input-paths" in run("-l", "x").stderr
    assert "unexpected argument" in run("-l", "x", "-i", "y", "--frobnicate").stderr


def test_validations_before_any_read_is_counted(example_dir, tmp_path):
    lib = os.path.join(example_dir, "library.fasta.gz")
    seq = os.path.join(example_dir, "sequence.fastq.gz")
    p = run("-l", lib, "-i", str(tmp_path / "missing.fq"))
    assert p.returncode == 1 and "Provided filepath does not exist" in p.stderr  # main.rs:133
    p = run("-l", lib, "-i", seq, seq, "-n", "only_one")
    assert "Must provide as many sample names as there are input files" in p.stderr  # main.rs:156
    bad = tmp_path / "ragged.fa"
    bad.write_text(">a\nACGT\n>b\nACG\n")
    p = run("-l", str(bad), "-i", seq)
    assert "Library sequence sizes are inconsistent" in p.stderr  # library.rs:83
    g = tmp_path / "g2s.txt"
    g.write_text("gene.0\tlib.0\n")
    p = run("-l", lib, "-i", seq, "-g", str(g))
    assert "Missing sgRNA aliases in gene map" in p.stderr  # count.rs:90-95
    g.write_text("gene.0 lib.0\n")
    assert "Missing '\\t' in gene map" in run("-l", lib, "-i", seq, "-g", str(g)).stderr  # genemap.rs:58
    g.write_text("gene.0\tlib.0\ngene.1\tlib.0\n")
    assert "Duplicate sgRNA key" in run("-l", lib, "-i", seq, "-g", str(g)).stderr  # genemap.rs:60
    short = tmp_path / "short.fq"
    short.write_text("@r\nACGTACGT\n+\nIIIIIIII\n")
    p = run("-l", lib, "-i", str(short))
    assert "Sequences in reference library are larger than the sequences in input" in p.stderr  # count.rs:98-100


# ---- whole runs (GPU) ------------------------------------------------------------------------

def table(text):
    lines = text.rstrip("\n").split("\n")
    return lines[0], sorted(lines[1:])


@pytest.mark.gpu
def test_example_run_matches_the_oracle_table(example_dir, expected, tmp_path):
    lib = os.path.join(example_dir, "library.fasta.gz")
    names = ["sequence", "zero.sequence", "diff.sequence", "offset", "offset_clipped"]
    paths = [os.path.join(example_dir, n + ".fastq.gz") for n in names]
    out = tmp_path / "counts.tsv"
    p = run("-l", lib, "-i", *paths, "-g", os.path.join(example_dir, "g2s.txt"), "-o", str(out), "-t", "3", check=True)
    assert "Calculated Offsets: [Forward(5), Forward(5), Forward(5), Forward(5), Forward(5)]" in p.stderr
    for n in names:
        fx = expected["fixtures"][n]
        frac = fx["matched_reads"] / fx["total_reads"]
        assert f"Finished: {n}; Fraction mapped: {frac:.3f} [{fx['matched_reads']} / {fx['total_reads']}]" in p.stderr

    olib_recs = orc.Records.from_path(lib)
    olib = orc.Library.from_reader(olib_recs)
    operm = orc.Permuter.new(olib)
    counters = [orc.Counter.new(orc.Records.from_path(x), olib, operm, orc.Offset.Forward(5)) for x in paths]
    g2s = open(os.path.join(example_dir, "g2s.txt"), "rb").read()
    want = orc.render_results(counters, names, olib, g2s, include_zero=False)
    assert table(out.read_text()) == table(want)
    first = out.read_text().split("\n")[1].split("\t")
    assert first[:2] == ["lib.0", "gene.0"] and int(first[2]) == expected["fixtures"]["sequence"]["counts"][0]


@pytest.mark.gpu
def test_flags_exact_offset_zero_rows_names_and_stdout(example_dir, expected, tmp_path):
    lib = os.path.join(example_dir, "library.fasta.gz")
    zero = os.path.join(example_dir, "zero.sequence.fastq.gz")
    p = run("-l", lib, "-i", zero, "-x", "-a", "5", "-q", "-n", "Z", check=True)
    assert p.stderr == ""
    head, rows = table(p.stdout)
    assert head == "Guide\tZ" and len(rows) == 90  # zero rows dropped (results.rs:90-94)
    p = run("-l", lib, "-i", zero, "-x", "-a", "5", "-q", "-z", check=True)
    head, rows = table(p.stdout)
    assert head == "Guide\tzero.sequence" and len(rows) == 100
    counts = {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows}
    assert [counts[f"lib.{i}"] for i in range(100)] == expected["fixtures"]["zero.sequence"]["counts"]
    # a wrong manual offset without recursion finds nothing; with recursion offset 4 finds all (Plus)
    p = run("-l", lib, "-i", zero, "-x", "-a", "4", "-p", "-q", "-z", check=True)
    assert all(r.endswith("\t0") for r in table(p.stdout)[1])
    p = run("-l", lib, "-i", zero, "-x", "-a", "4", "-q", check=True)
    assert len(table(p.stdout)[1]) == 90


@pytest.mark.gpu
def test_reverse_reads_plain_fasta_and_multi_member_gzip(example_dir, tmp_path):
    """reverse-complemented copies of the fixture, written as a two-member gzip FASTA"""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    seqs = [l for l in gzip.open(os.path.join(example_dir, "sequence.fastq.gz"), "rb").read().split(b"\n")[1::4]]
    rc = [s.translate(comp)[::-1] for s in seqs]
    path = tmp_path / "rev.fa.gz"
    half = len(rc) // 2
    with open(path, "wb") as f:
        for part in (rc[:half], rc[half:]):
            f.write(gzip.compress(b"".join(b">r\n" + s + b"\n" for s in part)))
    lib = os.path.join(example_dir, "library.fasta.gz")
    p = run("-l", lib, "-i", os.path.join(example_dir, "sequence.fastq.gz"), str(path), check=True)
    assert "Calculated Offsets: [Forward(5), Reverse(5)]" in p.stderr
    head, rows = table(p.stdout)
    assert head == "Guide\tsequence\trev"  # utils.rs:24-28 strips .gz then .fa
    assert all(r.split("\t")[1] == r.split("\t")[2] for r in rows) and len(rows) == 100


@pytest.mark.gpu
def test_gene_map_with_crlf_line_ends(example_dir, tmp_path):
    """genemap.rs:53-68 reads the map with bstr's for_byte_line, which strips "\\r\\n" as well as
    "\\n": a gene map saved on Windows must give the same table as the original."""
    lib = os.path.join(example_dir, "library.fasta.gz")
    fq = os.path.join(example_dir, "sequence.fastq.gz")
    g2s = os.path.join(example_dir, "g2s.txt")
    crlf = tmp_path / "g2s_crlf.txt"
    crlf.write_bytes(open(g2s, "rb").read().replace(b"\n", b"\r\n"))
    want = run("-l", lib, "-i", fq, "-g", g2s, "-q", check=True).stdout
    got = run("-l", lib, "-i", fq, "-g", str(crlf), "-q", check=True).stdout
    assert got == want and "\r" not in got
    # the oracle strips it too (oracle.cpp load_genemap): product and oracle agree
    lib_recs = orc.Records.from_path(lib)
    olib = orc.Library.from_reader(lib_recs)
    reads = orc.Records.from_path(fq)
    oc = orc.Counter.new(reads, olib, orc.Permuter.new(olib), orc.entropy_offset(lib_recs, reads, 5000))
    text = orc.render_results([oc], ["sequence"], olib, crlf.read_bytes(), include_zero=False)
    assert table(got) == table(text)


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [2, 5])
def test_one_sample_cut_into_read_shards_gives_the_same_table(shards, tmp_path):
    """--read-shards N: the batches of ONE sample are dealt to N counters and their vectors summed
    with sgc_reduce_counts (on one device: the fold; across devices: NCCL).  Same table, same
    totals, and the offset is detected once (offsetter.rs:192-200), not per shard."""
    from sgcount_b200 import synth

    seed = 0xB2000005
    arr = synth.make_library(seed, 5000, 20)
    lib_path = str(tmp_path / "lib.fa")
    with open(lib_path, "wb") as f:
        f.write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
    fq = str(tmp_path / "big.fastq.gz")
    # 3.5 M reads = 266 MB of sequence lines: five 64 MB batches, so every shard gets work
    synth.Sample(seed, 0, arr, 75, 11, True).write_fastq(fq, 0, 3_500_000, reads_per_member=400_000)
    one = run("-l", lib_path, "-i", fq, "-o", str(tmp_path / "one.tsv"), check=True)
    cut = run("-l", lib_path, "-i", fq, "-o", str(tmp_path / "cut.tsv"), "--read-shards", str(shards), "--timing",
              "--whole-lines", check=True)
    assert open(tmp_path / "one.tsv").read() == open(tmp_path / "cut.tsv").read()
    assert one.stderr.count("Calculated Offsets: [Reverse(11)]") == 1
    assert cut.stderr.count("Calculated Offsets: [Reverse(11)]") == 1
    fin = [l for l in one.stderr.splitlines() if l.startswith("Finished")]
    assert fin and fin == [l for l in cut.stderr.splitlines() if l.startswith("Finished")]
    # 266 MB of whole sequence lines are four 64 MB batches: at most four counters get work (as span
    # records, the default and what `one` used, the same reads are 84 MB)
    import re

    used = int(re.search(r'"read_shards_per_sample": (\d+)', cut.stderr).group(1))
    assert used == min(shards, 4)


@pytest.mark.gpu
def test_one_sample_over_every_device_with_nccl(tmp_path):
    """config 5's partitioning through the product: one sample, --gpus = every device of the box"""
    import torch

    from sgcount_b200 import synth

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two devices")
    seed = 0xB2000005
    arr = synth.make_library(seed, 20000, 20)
    lib_path = str(tmp_path / "lib.fa")
    with open(lib_path, "wb") as f:
        f.write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
    fq = str(tmp_path / "big.fastq.gz")
    n_reads = 900_000 * n_dev
    synth.Sample(seed, 0, arr, 75, 3, False).write_fastq(fq, 0, n_reads, reads_per_member=300_000)
    one = run("-l", lib_path, "-i", fq, "-o", str(tmp_path / "one.tsv"), check=True)
    # whole lines: 68 MB per device, so every device gets at least one 64 MB batch (as span
    # records, the default, the whole sample is one batch)
    many = run("-l", lib_path, "-i", fq, "-o", str(tmp_path / "many.tsv"), "--gpus", str(n_dev), "--read-shards", str(n_dev),
               "--timing", "--whole-lines", check=True)
    assert open(tmp_path / "one.tsv").read() == open(tmp_path / "many.tsv").read()
    assert '"read_shards_per_sample": %d' % n_dev in many.stderr
    assert many.stderr.count("Calculated Offsets: [Forward(3)]") == 1
    recs = orc.Records.from_path(fq)
    lib_recs = orc.Records.from_path(lib_path)
    olib = orc.Library.from_reader(lib_recs)
    oc = orc.Counter.new(recs, olib, orc.Permuter.new(olib), orc.Offset.Forward(3), None, True, n_threads=os.cpu_count() or 4)
    text = orc.render_results([oc], ["big"], olib, None, include_zero=False)
    assert table(open(tmp_path / "many.tsv").read()) == table(text)


@pytest.mark.gpu
def test_span_records_and_whole_lines_give_the_same_table(tmp_path):
    """fixed-length reads travel as span records by default; --whole-lines sends the lines.  A
    sample with one read of another length falls back member by member.  Same table each way."""
    from sgcount_b200 import synth

    seed = 0xB2000006
    arr = synth.make_library(seed, 3000, 20)
    lib_path = str(tmp_path / "lib.fa")
    with open(lib_path, "wb") as f:
        f.write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
    for rev, off in ((False, 0), (True, 17)):
        fq = str(tmp_path / f"s{int(rev)}.fastq.gz")
        synth.Sample(seed, int(rev), arr, 75, off, rev).write_fastq(fq, 0, 300_000, reads_per_member=40_000)
        a = run("-l", lib_path, "-i", fq, "--timing", check=True)
        b = run("-l", lib_path, "-i", fq, "--timing", "--whole-lines", check=True)
        assert a.stdout == b.stdout and len(a.stdout.splitlines()) > 1000
        assert '"span_reads": 300000' in a.stderr and '"span_reads": 0' in b.stderr
        # one shorter read in the last member: that member comes as whole lines
        text = gzip.open(fq, "rb").read().rstrip(b"\n").split(b"\n")
        text[-3] = text[-3][:60]
        text[-1] = text[-1][:60]
        odd = str(tmp_path / f"odd{int(rev)}.fastq.gz")
        members = [text[i:i + 160_000] for i in range(0, len(text), 160_000)]
        with open(odd, "wb") as f:
            for m in members:
                f.write(gzip.compress(b"\n".join(m) + b"\n", 1))
        c = run("-l", lib_path, "-i", odd, "-a", str(off), *(["-r"] if rev else []), "--timing", check=True)
        d = run("-l", lib_path, "-i", odd, "-a", str(off), *(["-r"] if rev else []), "--timing", "--whole-lines", check=True)
        assert c.stdout == d.stdout
        import re

        # the member with the short read comes as whole lines, and so does every member framed after
        # it was seen (the inflate threads run side by side: how many that is depends on timing)
        n_span = int(re.search(r'"span_reads": (\d+)', c.stderr).group(1))
        assert 0 <= n_span <= 300_000 - 20_000


@pytest.mark.gpu
def test_bgzf_samples_are_inflated_framed_and_counted_on_the_device(tmp_path):
    """BGZF input of fixed-length FASTQ takes the device ingest (sgc_fastq_stream_*); --host-inflate
    forces the host's inflate threads; a BGZF file the device declines (a read of another length)
    falls back by itself.  Always the same table, and the oracle's."""
    import json

    from sgcount_b200 import synth

    seed = 0xB2000007
    arr = synth.make_library(seed, 4000, 20)
    lib_path = str(tmp_path / "lib.fa")
    with open(lib_path, "wb") as f:
        f.write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
    paths = []
    for s, (rev, off) in enumerate([(False, 4), (True, 20)]):
        p = str(tmp_path / f"s{s}.fastq.gz")
        synth.Sample(seed, s, arr, 75, off, rev).write_fastq_bgzf(p, 0, 250_000)
        paths.append(p)

    def timing(proc):
        return json.loads([l for l in proc.stderr.splitlines() if l.startswith("{")][-1])

    dev = run("-l", lib_path, "-i", *paths, "--timing", check=True)
    host = run("-l", lib_path, "-i", *paths, "--timing", "--host-inflate", check=True)
    assert "Calculated Offsets: [Forward(4), Reverse(20)]" in dev.stderr
    assert dev.stdout == host.stdout and len(dev.stdout.splitlines()) > 1000
    assert timing(dev)["device_ingest_samples"] == 2 and timing(dev)["device_blocks"] > 1000
    assert timing(host)["device_ingest_samples"] == 0
    assert [l for l in dev.stderr.splitlines() if l.startswith("Finished")] == \
        [l for l in host.stderr.splitlines() if l.startswith("Finished")]
    # the oracle on the same files
    lib_recs = orc.Records.from_path(lib_path)
    olib = orc.Library.from_reader(lib_recs)
    operm = orc.Permuter.new(olib)
    counters = [orc.Counter.new(orc.Records.from_path(p), olib, operm, orc.Offset(rev, off), None, True,
                                n_threads=os.cpu_count() or 4) for p, (rev, off) in zip(paths, [(False, 4), (True, 20)])]
    text = orc.render_results(counters, ["s0", "s1"], olib, None, include_zero=False)
    assert table(dev.stdout) == table(text)
    # one shorter read deep in the file: the head looks regular, the device reports it, the host counts
    lines = gzip.open(paths[0], "rb").read().rstrip(b"\n").split(b"\n")
    lines[4 * 200_000 + 1] = lines[4 * 200_000 + 1][:50]
    lines[4 * 200_000 + 3] = lines[4 * 200_000 + 3][:50]
    odd_text = b"\n".join(lines) + b"\n"
    import struct
    import zlib

    odd = str(tmp_path / "odd.fastq.gz")
    with open(odd, "wb") as f:
        for i in range(0, len(odd_text), 0xff00):
            b = odd_text[i:i + 0xff00]
            raw = zlib.compressobj(1, zlib.DEFLATED, -15)
            body = raw.compress(b) + raw.flush()
            f.write(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0]) + b"BC" + struct.pack("<HH", 2, 18 + len(body) + 8 - 1) +
                    body + struct.pack("<II", zlib.crc32(b), len(b)))
    a = run("-l", lib_path, "-i", odd, "-a", "4", "--timing", check=True)
    b = run("-l", lib_path, "-i", odd, "-a", "4", "--timing", "--host-inflate", check=True)
    assert a.stdout == b.stdout
    # the span mode reports the odd read; the CLI starts over in the variable-length mode, still on the device
    assert timing(a)["device_ingest_samples"] == 1 and timing(a)["span_reads"] == 0
    # a FASTA sample in BGZF blocks is not framed on the device at all
    seqs = [l for l in lines[1::4]][:20000]
    fa = str(tmp_path / "reads.fa.gz")
    with open(fa, "wb") as f:
        text = b"".join(b">r\n" + s + b"\n" for s in seqs)
        for i in range(0, len(text), 0xff00):
            blk = text[i:i + 0xff00]
            raw = zlib.compressobj(1, zlib.DEFLATED, -15)
            body = raw.compress(blk) + raw.flush()
            f.write(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0]) + b"BC" + struct.pack("<HH", 2, 18 + len(body) + 8 - 1) +
                    body + struct.pack("<II", zlib.crc32(blk), len(blk)))
    c = run("-l", lib_path, "-i", fa, "-a", "4", "--timing", check=True)
    assert timing(c)["device_ingest_samples"] == 0 and timing(c)["host_ingest_because"] == "not FASTQ"
    assert len(c.stdout.splitlines()) > 500


@pytest.mark.gpu
@pytest.mark.parametrize("variable", [False, True])
def test_bgzf_sample_cut_into_independent_waves_over_several_counters(variable, tmp_path):
    """--read-shards N on BGZF input: the host cuts the blocks into waves that start and end on record
    boundaries (it inflates the one block at each cut), the waves go to N streams / counters (devices,
    when there are several) and the count vectors are summed.  Same table as one dependent stream."""
    import json
    import struct
    import zlib

    from sgcount_b200 import synth

    seed = 0xB2000008
    arr = synth.make_library(seed, 3000, 20)
    lib_path = str(tmp_path / "lib.fa")
    with open(lib_path, "wb") as f:
        f.write(b"".join(b">lib.%d\n%s\n" % (i, arr[i].tobytes()) for i in range(len(arr))))
    fq = str(tmp_path / "s.fastq.gz")
    if not variable:
        synth.Sample(seed, 0, arr, 75, 6, False).write_fastq_bgzf(fq, 0, 400_000)
    else:
        plain = str(tmp_path / "plain.fastq.gz")
        synth.Sample(seed, 0, arr, 75, 6, False).write_fastq(plain, 0, 200_000)
        lines = gzip.open(plain, "rb").read().rstrip(b"\n").split(b"\n")
        rng = __import__("random").Random(3)
        for r in range(0, len(lines), 4):  # trim a third of the reads (and their quality lines)
            if rng.random() < 0.33:
                n = rng.randrange(0, 70)
                lines[r + 1], lines[r + 3] = lines[r + 1][:n], lines[r + 3][:n]
        text = b"\n".join(lines) + b"\n"
        with open(fq, "wb") as f:
            for i in range(0, len(text), 0xff00):
                b = text[i:i + 0xff00]
                raw = zlib.compressobj(1, zlib.DEFLATED, -15)
                body = raw.compress(b) + raw.flush()
                f.write(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0]) + b"BC" + struct.pack("<HH", 2, 18 + len(body) + 8 - 1) +
                        body + struct.pack("<II", zlib.crc32(b), len(b)))
    one = run("-l", lib_path, "-i", fq, "-a", "6", "--timing", check=True)
    host = run("-l", lib_path, "-i", fq, "-a", "6", "--host-inflate", check=True)
    env = dict(os.environ, SGC_WAVE_BLOCKS="37")
    p = subprocess.run([BIN, "-l", lib_path, "-i", fq, "-a", "6", "--timing", "--read-shards", "3"], capture_output=True,
                       text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr
    t1 = json.loads([l for l in one.stderr.splitlines() if l.startswith("{")][-1])
    t3 = json.loads([l for l in p.stderr.splitlines() if l.startswith("{")][-1])
    assert t1["device_ingest_samples"] == 1 and t3["device_ingest_samples"] == 1
    assert t3["read_shards_per_sample"] == 3
    assert one.stdout == host.stdout == p.stdout and len(p.stdout.splitlines()) > 1000
    assert [l for l in one.stderr.splitlines() if l.startswith("Finished")] == [l for l in p.stderr.splitlines() if l.startswith("Finished")]
