"""The host's whole-member gzip decoder (sgcount_b200/host/inflate.cpp) against zlib: every kind
of DEFLATE block (stored, fixed, dynamic), long runs, far matches, second-level Huffman tables,
gzip header options, several members; and corrupt / truncated members must be DECLINED (the
caller then hands the bytes to zlib), never mis-decoded."""
import gzip
import os
import random
import struct
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "sgcount_b200", "lib", "fastx_dump")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(DUMP):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(DUMP)


def fnv(data: bytes) -> str:
    h = 1469598103934665603
    for c in data:
        h = ((h ^ c) * 1099511628211) & (2**64 - 1)
    return f"{len(data)} {h:x}"


def member(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, 31, memlevel, strategy)
    return c.compress(data) + c.flush()


def run(path):
    p = subprocess.run([DUMP, str(path), "1", "inflate"], capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout.strip()


def corpus():
    rng = random.Random(5)
    fastq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGT") for _ in range(75)), b"I" * 75) for i in range(3000))
    skewed = bytes(rng.choice(b"aaaaaaaabbbbccd\n") for _ in range(60000))
    many_symbols = bytes(int(rng.paretovariate(0.7)) % 256 for _ in range(80000))  # long and short codes together
    far = os.urandom(5000) + b"x" * 30000 + os.urandom(20) * 3 + b"y" * 2700
    return {"empty": b"", "one": b"A", "fastq": fastq, "zeros": bytes(100000), "random": os.urandom(70000),
            "skewed": skewed, "many_symbols": many_symbols, "far": far + far[:5000], "short_runs": b"ab" * 5000 + b"abc" * 3000}


@pytest.mark.parametrize("name", list(corpus()))
def test_matches_zlib_on_every_block_kind(tmp_path, name):
    data = corpus()[name]
    variants = {"l1": member(data, 1), "l6": member(data, 6), "l9": member(data, 9), "stored": member(data, 0),
                "fixed": member(data, 6, zlib.Z_FIXED), "huffman_only": member(data, 6, zlib.Z_HUFFMAN_ONLY),
                "rle": member(data, 6, zlib.Z_RLE), "mem1": member(data, 9, memlevel=1)}
    for vname, blob in variants.items():
        assert gzip.decompress(blob) == data
        path = tmp_path / f"{name}.{vname}.gz"
        path.write_bytes(blob)
        assert run(path) == (0, fnv(data)), (name, vname)


def test_header_options_and_several_members(tmp_path):
    data = b"@r\nACGT\n+\nIIII\n" * 700
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(data) + raw.flush()
    trailer = struct.pack("<II", zlib.crc32(data), len(data))
    # FEXTRA | FNAME | FCOMMENT | FHCRC
    header = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + struct.pack("<H", 5) + b"extra" + b"name.fq\0" + b"note\0"
    header += struct.pack("<H", zlib.crc32(header) & 0xFFFF)
    blob = header + body + trailer
    assert gzip.decompress(blob) == data
    (tmp_path / "opts.gz").write_bytes(blob + member(b"tail\n") + member(b""))
    assert run(tmp_path / "opts.gz") == (0, fnv(data + b"tail\n"))


def test_corrupt_and_truncated_members_are_declined(tmp_path):
    data = os.urandom(3000) + b"ACGT" * 5000
    blob = member(data, 6)
    cases = {"cut_body": blob[:len(blob) // 2], "cut_trailer": blob[:-3], "bad_crc": blob[:-8] + b"\0\0\0\0" + blob[-4:],
             "bad_isize": blob[:-4] + struct.pack("<I", len(data) + 1), "not_gzip": b"@r\nACGT\n+\nIIII\n" * 4,
             "bad_method": blob[:2] + b"\x07" + blob[3:]}
    flipped = bytearray(blob)
    flipped[len(blob) // 3] ^= 0x40
    cases["bit_flip"] = bytes(flipped)
    for name, b in cases.items():
        (tmp_path / name).write_bytes(b)
        assert run(tmp_path / name)[0] == 4, name


def test_reader_output_is_the_same_with_and_without_the_fast_decoder(tmp_path):
    rng = random.Random(9)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(rng.choice([75, 75, 80]))), b"F" * 75)
                    for i in range(40000))
    cuts = sorted(rng.sample(range(1, len(text)), 7))
    blob = b"".join(member(m, 1) for m in [text[a:b] for a, b in zip([0] + cuts, cuts + [len(text)])])
    (tmp_path / "x.fq.gz").write_bytes(blob)
    outs = set()
    for env in ({}, {"SGC_INFLATE": "zlib"}):
        p = subprocess.run([DUMP, str(tmp_path / "x.fq.gz"), "4", "blocks"], capture_output=True, text=True, timeout=120,
                           env={**os.environ, **env})
        assert p.returncode == 0, p.stderr
        outs.add(p.stdout.strip())
    assert len(outs) == 1 and outs.pop().startswith("40000 ")


def test_differential_fuzz_against_zlib(tmp_path):
    """Random structured inputs x random deflate parameters: the decoder must reproduce zlib's
    output; random corruptions of valid members must be declined or decode to exactly the original
    (a flip in the gzip header's ignored fields) — never to something else."""
    rng = random.Random(20240521)

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

    blobs, datas = [], []
    for case in range(60):
        data = b"".join(chunk() for _ in range(rng.randrange(1, 6)))
        level = rng.choice([0, 1, 1, 3, 6, 9])
        strategy = rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])
        c = zlib.compressobj(level, zlib.DEFLATED, 31, rng.choice([1, 5, 8, 9]), strategy)
        blob = b""
        at = 0
        while at < len(data):  # several blocks per member: flush at random points
            step = rng.randrange(1, len(data) + 1)
            blob += c.compress(data[at:at + step])
            if rng.random() < 0.5:
                blob += c.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))
            at += step
        blob += c.flush()
        assert gzip.decompress(blob) == data
        blobs.append(blob)
        datas.append(data)
        path = tmp_path / f"case{case}.gz"
        path.write_bytes(blob)
        assert run(path) == (0, fnv(data)), case
    # all of them as one multi-member file
    (tmp_path / "all.gz").write_bytes(b"".join(blobs))
    assert run(tmp_path / "all.gz") == (0, fnv(b"".join(datas)))
    # corruptions
    for case in range(80):
        i = rng.randrange(len(blobs))
        blob = bytearray(blobs[i])
        if len(blob) < 30:
            continue
        for _ in range(rng.choice([1, 1, 2, 5])):
            blob[rng.randrange(len(blob))] ^= 1 << rng.randrange(8)
        path = tmp_path / f"bad{case}.gz"
        path.write_bytes(bytes(blob))
        rc, out = run(path)
        assert rc == 4 or (rc == 0 and out == fnv(datas[i])), (case, rc, out)
