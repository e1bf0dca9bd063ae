"""The device gzip decoder's core (sgcount_b200/csrc/inflate_core.h), compiled for the HOST, against
zlib: the same source is what every thread of the device kernel runs (csrc/gzip.cu), so its
logic is checked here without a GPU — every kind of DEFLATE block, gzip header options, several
members, and corrupt / truncated members must come back as errors, never as other bytes.  The
corpus is the one the host decoder is tested with (tests/test_host_inflate.py)."""
import gzip
import os
import random
import struct
import subprocess
import zlib

import pytest

from test_host_inflate import corpus, fnv, member

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sgcount_b200", "lib", "inflate_core_test")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(EXE):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(EXE)


def run(path, cap=None, align=0):
    """decodes every member of the file; the output of member m is placed at alignment (m + align) % 8
    between guard bytes (on the device the neighbouring bytes belong to other threads)"""
    p = subprocess.run([EXE, str(path), str(cap or (1 << 30)), str(align)], capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout.strip()


def want(datas):
    total = b"".join(datas)
    return f"{len(datas)} {fnv(total)}"


@pytest.mark.parametrize("name", list(corpus()))
def test_matches_zlib_on_every_block_kind(tmp_path, name):
    data = corpus()[name]
    variants = {"l1": member(data, 1), "l6": member(data, 6), "l9": member(data, 9), "stored": member(data, 0),
                "fixed": member(data, 6, zlib.Z_FIXED), "huffman_only": member(data, 6, zlib.Z_HUFFMAN_ONLY),
                "rle": member(data, 6, zlib.Z_RLE), "mem1": member(data, 9, memlevel=1)}
    for vname, blob in variants.items():
        path = tmp_path / f"{name}.{vname}.gz"
        path.write_bytes(blob)
        for align in (0, 1, 5):
            assert run(path, align=align) == (0, want([data])), (name, vname, align)


def test_header_options_members_and_bgzf_blocks(tmp_path):
    data = b"@r\nACGT\n+\nIIII\n" * 700
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(data) + raw.flush()
    trailer = struct.pack("<II", zlib.crc32(data), len(data))
    header = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + struct.pack("<H", 5) + b"extra" + b"name.fq\0" + b"note\0"
    header += struct.pack("<H", zlib.crc32(header) & 0xFFFF)
    (tmp_path / "opts.gz").write_bytes(header + body + trailer + member(b"tail\n") + member(b""))
    assert run(tmp_path / "opts.gz") == (0, want([data, b"tail\n", b""]))
    # BGZF: 64 KB blocks with the BC extra field, and the empty end-of-file block
    rng = random.Random(3)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(75)), b"F" * 75) for i in range(5000))
    blocks = [text[i:i + 0xff00] for i in range(0, len(text), 0xff00)] + [b""]
    blob = b""
    for b in blocks:
        raw = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = raw.compress(b) + raw.flush()
        bsize = 12 + 6 + len(body) + 8 - 1
        blob += bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0]) + b"BC" + struct.pack("<HH", 2, bsize) + body + \
            struct.pack("<II", zlib.crc32(b), len(b))
    assert gzip.decompress(blob) == text
    (tmp_path / "bgzf.gz").write_bytes(blob)
    assert run(tmp_path / "bgzf.gz") == (0, want(blocks))


def test_corrupt_and_truncated_members_are_errors(tmp_path):
    data = os.urandom(3000) + b"ACGT" * 5000
    blob = member(data, 6)
    cases = {"cut_body": blob[:len(blob) // 2], "cut_trailer": blob[:-3], "not_gzip": b"@r\nACGT\n+\nIIII\n" * 4,
             "bad_method": blob[:2] + b"\x07" + blob[3:], "bad_isize": blob[:-4] + struct.pack("<I", len(data) + 1)}
    for name, b in cases.items():
        (tmp_path / name).write_bytes(b)
        assert run(tmp_path / name)[0] in (4, 5), name
    (tmp_path / "ok.gz").write_bytes(blob)
    assert run(tmp_path / "ok.gz", cap=len(data) - 1)[0] == 4  # output capacity is respected
    assert run(tmp_path / "ok.gz", cap=len(data)) == (0, want([data]))


def test_differential_fuzz_against_zlib(tmp_path):
    """random structured inputs x random deflate parameters; bit flips in valid members must be
    errors or decode to exactly what zlib makes of the same bytes"""
    rng = random.Random(77)

    def chunk():
        kind = rng.randrange(6)
        n = rng.choice([1, 2, 7, 50, 300, 5000, 40000])
        if kind == 0:
            return os.urandom(n)
        if kind == 1:
            return bytes([rng.randrange(256)]) * n
        if kind == 2:
            alphabet = bytes(rng.sample(range(256), rng.choice([2, 4, 20])))
            return bytes(rng.choice(alphabet) for _ in range(n))
        if kind == 3:
            unit = os.urandom(rng.choice([2, 3, 5, 9, 40]))
            return (unit * (n // len(unit) + 1))[:n]
        if kind == 4:
            return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(40)), b"F" * 40) for i in range(n // 90 + 1))
        return bytes(min(255, int(rng.expovariate(0.05))) for _ in range(n))

    for case in range(60):
        datas = [b"".join(chunk() for _ in range(rng.randrange(1, 4))) for _ in range(rng.randrange(1, 4))]
        blob = b""
        for d in datas:
            level = rng.choice([0, 1, 1, 3, 6, 9])
            strategy = rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])
            c = zlib.compressobj(level, zlib.DEFLATED, 31, rng.choice([1, 5, 8, 9]), strategy)
            blob += c.compress(d) + c.flush()
        path = tmp_path / f"case{case}.gz"
        path.write_bytes(blob)
        assert run(path, align=case) == (0, want(datas)), case
        # corrupt one bit of the first member's body
        if len(blob) > 40:
            bad = bytearray(blob)
            at = rng.randrange(12, len(blob) - 8)
            bad[at] ^= 1 << rng.randrange(8)
            path.write_bytes(bytes(bad))
            rc, out = run(path)
            try:
                ref = gzip.decompress(bytes(bad))  # zlib checks the CRC; the core reports size only
            except Exception:
                ref = None
            if rc == 0 and ref is not None:
                assert out.split()[1:] == fnv(ref).split(), case
