"""FASTQ straight from BGZF blocks (sgc_fastq_stream_*, csrc/gzip.cu): inflate, record framing and
counting on the device must give the table the host path gives, whatever the block boundaries,
wave sizes and compression parameters; irregular input must be REPORTED (the caller then counts
the sample through the host path), never miscounted."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import sgcount_b200 as sg
from sgcount_b200 import _cabi, synth

from helpers import make_library, make_reads

pytestmark = pytest.mark.gpu


def bgzf(text: bytes, block=0xff00, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, eof=True) -> bytes:
    out = b""
    parts = [text[i:i + block] for i in range(0, len(text), block)] + ([b""] if eof else [])
    for b in parts:
        raw = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        body = raw.compress(b) + raw.flush()
        out += bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0]) + b"BC" + struct.pack("<HH", 2, 18 + len(body) + 8 - 1) + \
            body + struct.pack("<II", zlib.crc32(b), len(b))
    return out


def fastq_text(seqs, final_newline=True):
    text = b"".join(b"@r%d some description\n%s\n+\n%s\n" % (i, s, b"F" * len(s)) for i, s in enumerate(seqs))
    return text if final_newline else text[:-1]


def host_counts(library, permuter, seqs, off, recursion=True):
    c = sg.Counter(library, permuter, off, recursion)
    c.submit(sg.ReadBatch.from_seqs(seqs))
    return c.finish()


def device_counts(library, permuter, blob, read_len, k, off, recursion=True, wave=64):
    """read_len > 0: span mode (fixed-length reads); read_len == 0: reads of any length"""
    if read_len:
        start, length, span_off = sg.span_geometry(k, read_len, off, recursion)
    else:
        start, length, span_off = 0, 0, off
    c = sg.Counter(library, permuter, span_off, recursion)
    stream = sg.FastqStream(c, read_len, start, length)
    begin, isize = sg.bgzf_blocks(blob)
    arr = np.frombuffer(blob, dtype=np.uint8)
    for a in range(0, len(isize), wave):
        b = min(len(isize), a + wave)
        stream.submit(arr, begin[a:b + 1], isize[a:b])
    n = stream.finish()
    return n, c.finish()


@pytest.mark.parametrize("reverse,offset,k", [(False, 5, 20), (True, 12, 20), (False, 0, 16), (True, 31, 24)])
def test_device_path_equals_host_path(reverse, offset, k):
    rng = np.random.default_rng(31 + offset)
    guides = make_library(rng, 500, k)
    read_len = 75
    seqs = make_reads(rng, guides, 30_000, read_len, offset, reverse, False, wild=b"J" if reverse else b"N")
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    off = sg.Offset(reverse, offset)
    want = host_counts(library, permuter, seqs, off)
    text = fastq_text(seqs)
    for label, blob, wave in [("l6", bgzf(text), 64), ("l1 small waves", bgzf(text, level=1), 3),
                              ("l9 tiny blocks", bgzf(text, block=777, level=9), 50),
                              ("fixed codes", bgzf(text, strategy=zlib.Z_FIXED), 1000),
                              ("stored", bgzf(text, level=0), 7), ("no eof block", bgzf(text, eof=False), 11),
                              ("no final newline", bgzf(fastq_text(seqs, False)), 5)]:
        assert gzip.decompress(blob).rstrip(b"\n") == text.rstrip(b"\n")
        n, got = device_counts(library, permuter, blob, read_len, k, off, wave=wave)
        assert n == len(seqs), label
        assert np.array_equal(got[0], want[0]) and got[1:] == want[1:], label


@pytest.mark.parametrize("reverse,offset", [(False, 5), (True, 9)])
def test_reads_of_any_length_are_counted_in_place(reverse, offset):
    """variable-length mode (read_len = 0): adapter-trimmed reads, some shorter than offset + k, their
    sequence lines counted where they lie in the inflated text by the line kernel"""
    rng = np.random.default_rng(52 + offset)
    k = 20
    guides = make_library(rng, 400, k)
    seqs = make_reads(rng, guides, 20_000, 75, offset, reverse, True, wild=b"J" if reverse else b"N")
    assert len({len(s) for s in seqs}) > 5
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    off = sg.Offset(reverse, offset)
    want = host_counts(library, permuter, seqs, off)
    text = fastq_text(seqs)
    for label, blob, wave in [("l6", bgzf(text), 64), ("l1 small waves", bgzf(text, level=1), 2),
                              ("tiny blocks", bgzf(text, block=999, level=9), 40), ("no final newline", bgzf(fastq_text(seqs, False)), 9)]:
        n, got = device_counts(library, permuter, blob, 0, k, off, wave=wave)
        assert n == len(seqs), label
        assert np.array_equal(got[0], want[0]) and got[1:] == want[1:], label
    # fixed-length reads go through this mode just as well
    fixed = make_reads(rng, guides, 5000, 60, offset, reverse, False)
    n, got = device_counts(library, permuter, bgzf(fastq_text(fixed)), 0, k, off)
    want = host_counts(library, permuter, fixed, off)
    assert n == 5000 and np.array_equal(got[0], want[0]) and got[1:] == want[1:]


def test_synthetic_bgzf_file_and_full_size_blocks(tmp_path):
    """the generator's BGZF writer (64 KB blocks cut anywhere, like bgzip), 1 M reads, Brunello-sized library"""
    seed = 0xB2000002
    arr = synth.make_library(seed, 77441, 20)
    library = sg.Library([arr[i].tobytes() for i in range(len(arr))], [b"lib.%d" % i for i in range(len(arr))])
    permuter = sg.Permuter.new(library)
    sample = synth.Sample(seed, 0, arr, 75, 5, False)
    n = 1_000_000
    path = str(tmp_path / "s.fastq.gz")
    sample.write_fastq_bgzf(path, 0, n)
    lines = sample.fill_host(0, n)
    want = sg.Counter(library, permuter, sg.Offset.Forward(5))
    want.submit(sg.ReadBatch(lines, n, None, 76, 75))
    blob = open(path, "rb").read()
    got_n, got = device_counts(library, permuter, blob, 75, 20, sg.Offset.Forward(5), wave=2000)
    w = want.finish()
    assert got_n == n and np.array_equal(got[0], w[0]) and got[1:] == w[1:]


def test_irregular_input_is_reported_not_miscounted():
    rng = np.random.default_rng(8)
    guides = make_library(rng, 100, 20)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    seqs = make_reads(rng, guides, 5000, 75, 5)
    off = sg.Offset.Forward(5)
    # a read of another length
    odd = list(seqs)
    odd[3777] = odd[3777][:60]
    with pytest.raises(sg.SgcError) as e:
        device_counts(library, permuter, bgzf(fastq_text(odd)), 75, 20, off)
    assert e.value.code == _cabi.ERR_FASTQ_FORMAT
    # FASTA, not FASTQ
    fasta = b"".join(b">r\n%s\n" % s for s in seqs)
    with pytest.raises(sg.SgcError) as e:
        device_counts(library, permuter, bgzf(fasta), 75, 20, off)
    assert e.value.code == _cabi.ERR_FASTQ_FORMAT
    # a truncated last record
    with pytest.raises(sg.SgcError) as e:
        device_counts(library, permuter, bgzf(fastq_text(seqs)[:-100]), 75, 20, off)
    assert e.value.code == _cabi.ERR_FASTQ_FORMAT
    # a corrupt block
    blob = bytearray(bgzf(fastq_text(seqs)))
    begin, _ = sg.bgzf_blocks(bytes(blob))
    blob[int(begin[2]) + 40] ^= 0x55
    with pytest.raises(sg.SgcError) as e:
        device_counts(library, permuter, bytes(blob), 75, 20, off)
    assert e.value.code in (_cabi.ERR_GZIP, _cabi.ERR_FASTQ_FORMAT)
    # a damaged byte that leaves the DEFLATE structure intact (a stored block): only the CRC-32 sees it
    stored = bytearray(bgzf(fastq_text(seqs), level=0))
    begin, _ = sg.bgzf_blocks(bytes(stored))
    at = int(begin[1]) + 18 + 5 + 1000  # inside the second block's literal bytes, in a quality line or a header
    stored[at] ^= 0x01
    with pytest.raises(sg.SgcError) as e:
        device_counts(library, permuter, bytes(stored), 75, 20, off)
    assert e.value.code in (_cabi.ERR_GZIP, _cabi.ERR_FASTQ_FORMAT)
    if e.value.code == _cabi.ERR_GZIP:
        assert "CRC-32" in str(e.value)
    # and the plain case still works afterwards
    n, got = device_counts(library, permuter, bgzf(fastq_text(seqs)), 75, 20, off)
    want = host_counts(library, permuter, seqs, off)
    assert n == 5000 and np.array_equal(got[0], want[0])
