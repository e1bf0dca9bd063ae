"""The host's FASTA/FASTQ reader (sgcount_b200/host/fastx.cpp), the stand-in for the `fxread`
crate: plain, gzip and multi-member gzip inflated by a thread pool must yield the same records
as Python's own parsing; malformed input must be reported, not mis-parsed."""
import gzip
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "sgcount_b200", "lib", "fastx_dump")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(DUMP):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(DUMP)


def fnv(records, with_ids=True):
    h = 1469598103934665603
    for rid, seq in records:
        for c in seq:
            h = ((h ^ c) * 1099511628211) & (2**64 - 1)
        h = ((h ^ 0xFF) * 1099511628211) & (2**64 - 1)
        if with_ids:
            for c in rid:
                h = ((h ^ c) * 1099511628211) & (2**64 - 1)
    return f"{len(records)} {h:x}"


def dump(path, threads, *extra):
    p = subprocess.run([DUMP, str(path), str(threads), *extra], capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout.strip(), p.stderr


def make_records(n, seed, fastq=True):
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        length = rng.choice([75, 75, 75, 80, 12, 0])
        recs.append((b"r%d extra" % i, bytes(rng.choice(b"ACGTNacgt") for _ in range(length))))
    recs[3] = (b"\x1f\x8b\x08\x00 looks like a gzip header", b"ACGT")
    if fastq:
        text = b"".join(b"@" + i + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in recs)
    else:
        text = b"".join(b">" + i + b"\n" + s + b"\n" for i, s in recs)
    return recs, text


@pytest.mark.parametrize("fastq", [True, False])
def test_plain_gzip_and_multi_member_agree(tmp_path, fastq):
    recs, text = make_records(30000, 7 + fastq, fastq)
    want = fnv(recs)
    rng = random.Random(1)
    cuts = sorted(rng.sample(range(1, len(text)), 9))  # member boundaries fall inside records
    members = [text[a:b] for a, b in zip([0] + cuts, cuts + [len(text)])]
    files = {"plain.fx": text, "single.fx.gz": gzip.compress(text, 1),
             "multi.fx.gz": b"".join(gzip.compress(m, 1) for m in members),
             "with_empty_member.fx.gz": gzip.compress(text[:100], 1) + gzip.compress(b"", 1) + gzip.compress(text[100:], 1)}
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
        for threads in (1, 3, 16):
            rc, out, err = dump(tmp_path / name, threads)
            assert (rc, out) == (0, want), (name, threads, err)
    rc, out, _ = dump(tmp_path / "multi.fx.gz", 4, "seq")
    assert (rc, out) == (0, fnv(recs, with_ids=False))


def test_last_line_without_newline_and_empty_file(tmp_path):
    (tmp_path / "a.fa").write_bytes(b">x\nACGT\n>y\nGG")
    assert dump(tmp_path / "a.fa", 1)[1] == fnv([(b"x", b"ACGT"), (b"y", b"GG")])
    (tmp_path / "empty.fq").write_bytes(b"")
    assert dump(tmp_path / "empty.fq", 1)[:2] == (0, fnv([]))


def test_malformed_inputs_are_errors(tmp_path):
    (tmp_path / "bad.fq").write_bytes(b"ACGT\nACGT\n")
    assert dump(tmp_path / "bad.fq", 1)[0] == 1
    (tmp_path / "trunc.fq").write_bytes(b"@r\nACGT\n+\n")
    assert "truncated" in dump(tmp_path / "trunc.fq", 1)[2]
    (tmp_path / "trunc.fa").write_bytes(b">r\n")
    assert "truncated" in dump(tmp_path / "trunc.fa", 1)[2]
    blob = gzip.compress(b"@r\nACGT\n+\nIIII\n" * 5000, 1)
    (tmp_path / "cut.fq.gz").write_bytes(blob[:len(blob) // 2])
    assert dump(tmp_path / "cut.fq.gz", 1)[0] == 1
    (tmp_path / "cut2.fq.gz").write_bytes(gzip.compress(b"@a\nAC\n+\nII\n", 1) * 3 + blob[:len(blob) // 2])
    assert dump(tmp_path / "cut2.fq.gz", 4)[0] == 1
    (tmp_path / "notgz.fq.gz").write_bytes(b"@r\nACGT\n+\nIIII\n")
    assert dump(tmp_path / "notgz.fq.gz", 2)[0] == 1


# ---- SeqBlockReader: the packed-sequence-line path of count_sample ----------------------------
@pytest.mark.parametrize("fastq", [True, False])
def test_blocks_agree_with_record_reader(tmp_path, fastq):
    """Members cut inside records (the workers' record-boundary assumption fails and the consumer
    re-frames), members cut ON record boundaries (the workers' blocks are used as they are), an
    empty member, plain and single-member input: always the sequences of the record reader."""
    recs, text = make_records(30000, 11 + fastq, fastq)
    want = fnv(recs, with_ids=False)
    rng = random.Random(2)
    cuts = sorted(rng.sample(range(1, len(text)), 9))
    inside = [text[a:b] for a, b in zip([0] + cuts, cuts + [len(text)])]
    starts = [m.start() for m in __import__("re").finditer(rb"^[@>]r\d+ extra$", text, flags=__import__("re").M)]
    bounds = sorted(rng.sample(starts[1:], 9))
    aligned = [text[a:b] for a, b in zip([0] + bounds, bounds + [len(text)])]
    mixed = aligned[:4] + [aligned[4][:50], aligned[4][50:]] + aligned[5:]  # one boundary inside a record
    files = {"plain.fx": text, "single.fx.gz": gzip.compress(text, 1),
             "inside.fx.gz": b"".join(gzip.compress(m, 1) for m in inside),
             "aligned.fx.gz": b"".join(gzip.compress(m, 1) for m in aligned),
             "mixed.fx.gz": b"".join(gzip.compress(m, 1) for m in mixed),
             "with_empty_member.fx.gz": b"".join(gzip.compress(m, 1) for m in aligned[:3] + [b""] + aligned[3:])}
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
        for threads in (1, 3, 16):
            rc, out, err = dump(tmp_path / name, threads, "blocks")
            assert (rc, out) == (0, want), (name, threads, err)


def test_blocks_edge_cases(tmp_path):
    (tmp_path / "a.fa").write_bytes(b">x\nACGT\n>y\nGG")
    assert dump(tmp_path / "a.fa", 1, "blocks")[1] == fnv([(b"x", b"ACGT"), (b"y", b"GG")], with_ids=False)
    (tmp_path / "empty.fq").write_bytes(b"")
    assert dump(tmp_path / "empty.fq", 1, "blocks")[:2] == (0, fnv([]))
    # last member ends without a newline, several members, several threads
    big = b"@r\nACGT\n+\nIIII\n" * 6000
    blob = gzip.compress(big, 1) + gzip.compress(big, 1) + gzip.compress(b"@z\nAC\n+\nII", 1)
    (tmp_path / "tail.fq.gz").write_bytes(blob)
    assert dump(tmp_path / "tail.fq.gz", 4, "blocks")[1] == fnv([(b"", b"ACGT")] * 12000 + [(b"", b"AC")], with_ids=False)
    (tmp_path / "bad.fq").write_bytes(b"ACGT\nACGT\n")
    assert dump(tmp_path / "bad.fq", 1, "blocks")[0] == 1
    (tmp_path / "trunc.fq").write_bytes(b"@r\nACGT\n+\n")
    assert "truncated" in dump(tmp_path / "trunc.fq", 1, "blocks")[2]
    (tmp_path / "trunc.fq.gz").write_bytes(gzip.compress(big, 1) + gzip.compress(big + b"@r\nACGT\n", 1))
    assert "truncated" in dump(tmp_path / "trunc.fq.gz", 4, "blocks")[2]
    (tmp_path / "cut.fq.gz").write_bytes(blob[:len(blob) // 3])
    assert dump(tmp_path / "cut.fq.gz", 4, "blocks")[0] == 1


def span_fnv(records, read_len, start, length):
    h = 1469598103934665603
    for _, seq in records:
        part, mark = (seq[start:start + length], 0xFE) if len(seq) == read_len else (seq, 0xFF)
        for c in part:
            h = ((h ^ c) * 1099511628211) & (2**64 - 1)
        h = ((h ^ mark) * 1099511628211) & (2**64 - 1)
    return f"{len(records)} {h:x}"


def test_span_framing(tmp_path):
    """Span records (fastx.h SpanSpec): bytes [start, start + len) of every read, fixed stride, framed
    by the inflate threads while every read has the sample's read length; a member holding a read
    of another length comes as whole lines, and so does everything framed by the consumer itself
    (plain files, single members, members cut inside a record).  What a counter looks at is the
    same either way."""
    rng = random.Random(5)
    recs = [(b"r%d extra" % i, bytes(rng.choice(b"ACGTN") for _ in range(75))) for i in range(40000)]
    text = b"".join(b"@" + i + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in recs)
    starts = [m.start() for m in __import__("re").finditer(rb"^@r\d+ extra$", text, flags=__import__("re").M)]
    bounds = sorted(rng.sample(starts[1:], 7))
    aligned = [text[a:b] for a, b in zip([0] + bounds, bounds + [len(text)])]
    spec = "75,4,22,24"
    want = span_fnv(recs, 75, 4, 22)
    (tmp_path / "aligned.fq.gz").write_bytes(b"".join(gzip.compress(m, 1) for m in aligned))
    for threads in (2, 16):
        rc, out, err = dump(tmp_path / "aligned.fq.gz", threads, "spans", spec)
        assert rc == 0 and out == want + " 40000", (threads, out, err)  # every record travelled as a span
    # framed by the consumer: whole lines, same content
    (tmp_path / "plain.fq").write_bytes(text)
    (tmp_path / "single.fq.gz").write_bytes(gzip.compress(text, 1))
    for name in ("plain.fq", "single.fq.gz"):
        rc, out, err = dump(tmp_path / name, 4, "spans", spec)
        assert rc == 0 and out == want + " 0", (name, out, err)
    # one read of another length in the middle of member 3: that member (at least) falls back
    odd = list(recs)
    k = sum(m.count(b"\n@r") + (1 if m.startswith(b"@r") else 0) for m in aligned[:3]) + 5
    odd[k] = (odd[k][0], odd[k][1][:40])
    text2 = b"".join(b"@" + i + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in odd)
    starts2 = [m.start() for m in __import__("re").finditer(rb"^@r\d+ extra$", text2, flags=__import__("re").M)]
    idx = [starts.index(b) for b in bounds]
    aligned2 = [text2[a:b] for a, b in zip([0] + [starts2[i] for i in idx], [starts2[i] for i in idx] + [len(text2)])]
    (tmp_path / "odd.fq.gz").write_bytes(b"".join(gzip.compress(m, 1) for m in aligned2))
    for threads in (2, 16):
        rc, out, err = dump(tmp_path / "odd.fq.gz", threads, "spans", spec)
        n, h, n_span = out.split()
        assert rc == 0 and f"{n} {h}" == span_fnv(odd, 75, 4, 22), (threads, out, err)
        assert int(n_span) <= 40000 - len([1 for _ in aligned2[3].split(b"\n+\n")]) + 1


def test_last_record_end_of_a_text_that_starts_anywhere(tmp_path):
    """fastq_last_record_end: where a BGZF block's text can be cut so that the next wave of the device
    ingest starts on a record — the text begins anywhere inside a record, quality lines may begin
    with '@' or '+', reads have any length"""
    rng = random.Random(14)
    recs = []
    for i in range(400):
        n = rng.choice([75, 75, 60, 1, 0])
        qual = bytes(rng.choice(b"@+IF#") for _ in range(n))
        recs.append(b"@r%d x\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(n)), qual))
    text = b"".join(recs)
    ends, at = [], 0
    for r in recs:
        at += len(r)
        ends.append(at)
    for trial in range(60):
        a = rng.randrange(0, len(text) - 2000)
        b = rng.randrange(a + 700, min(len(text), a + 5000))
        (tmp_path / "piece").write_bytes(text[a:b])
        rc, out, _ = dump(tmp_path / "piece", 1, "cut")
        # the last record wholly inside [a, b) that does not start at a itself (its start cannot be told)
        inside = [e for s, e in zip([0] + ends, ends) if s > a and e <= b]
        assert rc == 0 and out == (str(inside[-1] - a) if inside else "none"), (trial, a, b)
    (tmp_path / "piece").write_bytes(b"ACGT\nACGT")
    assert dump(tmp_path / "piece", 1, "cut")[1] == "none"


def test_first_record_start_of_a_text_that_starts_anywhere(tmp_path):
    """fastx_first_record_start: where an inflate thread starts framing a member (a run of BGZF
    blocks) whose text begins anywhere inside a record; quality lines may begin with '@' or '+'."""
    rng = random.Random(15)
    recs = []
    for i in range(300):
        n = rng.choice([75, 75, 60, 1])
        qual = bytes(rng.choice(b"@+IF#") for _ in range(n))
        recs.append(b"@r%d x\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(n)), qual))
    text = b"".join(recs)
    starts, at = [], 0
    for r in recs:
        starts.append(at)
        at += len(r)
    for trial in range(80):
        a = rng.choice(starts) if trial % 4 == 0 else rng.randrange(0, len(text) - 3000)
        (tmp_path / "piece").write_bytes(text[a:a + 2500])
        rc, out, _ = dump(tmp_path / "piece", 1, "start", "4")
        first = min(s for s in starts if s >= a)
        # a text that begins inside a header line can look like a record start itself when the rest of the
        # header begins with '@' (none here); otherwise the answer is the first true record start
        assert rc == 0 and out == str(first - a), (trial, a, out)
    fa = b"".join(b">g%d\n%s\n" % (i, b"ACGT" * 5) for i in range(50))
    for a in (0, 1, 5, 9, 30):
        (tmp_path / "piece").write_bytes(fa[a:])
        want = fa.index(b"\n>", a - 1 if a else 0) + 1 - a if a and fa[a:a + 1] != b">" else 0
        assert dump(tmp_path / "piece", 1, "start", "2")[1] == str(want), a
    (tmp_path / "piece").write_bytes(b"ACGT\nACGT")
    assert dump(tmp_path / "piece", 1, "start", "4")[1] == "none"
    assert dump(tmp_path / "piece", 1, "start", "2")[1] == "none"


def write_bgzf(path, text, block=65280, level=1):
    import struct
    import zlib

    with open(path, "wb") as f:
        for at in list(range(0, len(text), block)) + [None]:  # the last one is the empty end-of-file block
            chunk = b"" if at is None else text[at:at + block]
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            d = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(d) + 25) + d
                    + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


@pytest.mark.parametrize("kind", ["fastq", "fastq_ragged", "fastq_crlf", "fasta"])
def test_bgzf_blocks_are_framed_by_the_inflate_threads(tmp_path, kind):
    """BGZF cuts the text every 64 KB, inside records: the inflate threads take runs of blocks, frame
    each run from the first record start they recognise, and the consumer only finishes the record
    before it.  Same sequences as the sequential reader, and (uniform FASTQ, FASTA) no run framed twice."""
    rng = random.Random(21)
    n = 60000
    nl = b"\r\n" if kind == "fastq_crlf" else b"\n"
    recs = []
    for i in range(n):
        length = rng.choice([75, 75, 75, 80, 12, 0]) if kind == "fastq_ragged" else 75
        recs.append((b"r%d x" % i, bytes(rng.choice(b"ACGTN") for _ in range(length))))
    if kind == "fasta":
        text = b"".join(b">" + i + nl + s + nl for i, s in recs)
    else:
        text = b"".join(b"@" + i + nl + s + nl + b"+" + nl + bytes(rng.choice(b"@+IF#") for _ in s) + nl for i, s in recs)
    if kind == "fastq_crlf":
        recs = [(i, s + b"\r") for i, s in recs]  # the reader hands the line over as it is; the CLI strips the '\r'
    path = tmp_path / "reads.fx.gz"
    write_bgzf(path, text)
    want = fnv(recs, with_ids=False)
    for threads in (1, 2, 8):
        rc, out, err = dump(path, threads, "blocks")
        assert (rc, out) == (0, want), (threads, err)
        if threads > 1:
            adopted, reframed = (int(x) for x in err.split()[1::2])
            assert adopted >= 4, err
            assert reframed == 0 or kind == "fastq_ragged", err
    if kind == "fastq":
        spec = (75, 10, 22, 24)
        rc, out, err = dump(path, 8, "spans", ",".join(map(str, spec)))
        n_out, h, n_span = out.split()
        assert rc == 0 and f"{n_out} {h}" == span_fnv(recs, 75, 10, 22) and int(n_span) == n, (out, err)
        assert err.split()[3] == "0", err
    # a block that is not what its header says: the chain breaks, the rest is read sequentially or reported
    blob = bytearray(path.read_bytes())
    blob[len(blob) // 2] ^= 0x55
    (tmp_path / "bad.fx.gz").write_bytes(bytes(blob))
    assert dump(tmp_path / "bad.fx.gz", 8, "blocks")[0] == 1


def test_bgzf_index_walked_in_parts_equals_the_sequential_walk(tmp_path):
    """bgzf_index on a large file walks the blocks in parts, side by side, each part from a guessed
    block start that the part before it must arrive at.  Same index as the sequential walk for every
    thread count; a planted look-alike (three chained block headers inside one block's stored bytes,
    right after a cut) is not confirmed and the file is walked sequentially."""
    import struct
    import zlib

    rng = random.Random(5)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGT") for _ in range(75)), b"I" * 75) for i in range(40000))
    path = tmp_path / "reads.fq.gz"
    write_bgzf(path, text, block=4000)
    rc, want, _ = dump(path, 1, "bgzfindex", "0")
    n_blocks, _, in_parts = want.split()
    assert rc == 0 and int(n_blocks) == len(text) // 4000 + 2 and in_parts == "0"
    for threads in (2, 3, 8, 16):
        rc, out, _ = dump(path, threads, "bgzfindex", "0")
        assert rc == 0 and out.split()[:2] == want.split()[:2] and out.split()[2] == "1", (threads, out)
    # below the size limit the walk stays sequential
    assert dump(path, 8, "bgzfindex", str(1 << 30))[1] == want

    # a file whose second half begins, inside a STORED block, with bytes that read as three chained
    # empty BGZF blocks: the guess of part 2 lands on them, the walk of part 1 steps over them
    def member(payload_deflate, crc, isize):
        return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(payload_deflate) + 25) + payload_deflate
                + struct.pack("<II", crc, isize))

    empty = member(b"\x03\0", 0, 0)
    fake = empty * 3
    filler = bytes(rng.choice(b"ACGT") for _ in range(3000))
    payload = filler + fake + filler  # one stored block: 01 len ~len bytes
    stored = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xFFFF) + payload
    tricky = member(stored, zlib.crc32(payload), len(payload))
    one = member(zlib.compress(filler, 1)[2:-4], zlib.crc32(filler), len(filler))
    head = one + one
    blob = head + tricky + one + empty
    # the cut of a two-part walk (len / 2) must fall in the filler just before the look-alike
    at_fake = len(head) + 18 + 5 + len(filler)
    assert at_fake - 2000 < len(blob) // 2 <= at_fake, (at_fake, len(blob))
    (tmp_path / "tricky.gz").write_bytes(blob)
    seq = dump(tmp_path / "tricky.gz", 1, "bgzfindex", "0")[1]
    par = dump(tmp_path / "tricky.gz", 2, "bgzfindex", "0")[1]
    assert seq.split()[0] == "5" and par.split()[:2] == seq.split()[:2] and par.split()[2] == "0", (seq, par)
    # not BGZF at all
    (tmp_path / "plain.gz").write_bytes(gzip.compress(text[:100000]))
    assert dump(tmp_path / "plain.gz", 8, "bgzfindex", "0")[1] == "not bgzf"


def test_bgzf_index_in_parts_on_random_files(tmp_path):
    """Random BGZF files — stored and compressed blocks of every size, payloads with chained block
    headers planted in them, now and then trailing bytes that are no block — give the same index
    (or the same refusal) whatever the number of parts."""
    import struct
    import zlib

    def member(payload, level):
        if level < 0:  # a stored block keeps planted bytes as they are
            body = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xFFFF) + payload
        else:
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = c.compress(payload) + c.flush()
        return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25) + body
                + struct.pack("<II", zlib.crc32(payload), len(payload)))

    fake = member(b"", 1)
    for seed in range(40):
        rng = random.Random(seed)
        blocks = []
        for _ in range(rng.randint(1, 40)):
            payload = bytes(rng.choice(b"ACGT\n@+I") for _ in range(rng.choice([0, 10, 500, 3000, 20000])))
            if rng.random() < 0.4:
                at = rng.randint(0, len(payload))
                payload = payload[:at] + fake * rng.randint(1, 4) + payload[at:]
            blocks.append(member(payload, rng.choice([-1, -1, 1, 6])))
        blob = b"".join(blocks) + (b"garbage" if rng.random() < 0.1 else b"")
        path = tmp_path / "f.gz"
        path.write_bytes(blob)
        want = dump(path, 1, "bgzfindex", "0")[1].split()[:2]
        for threads in (2, 5, 16):
            assert dump(path, threads, "bgzfindex", "0")[1].split()[:2] == want, (seed, threads)
