import os, random, struct, subprocess, zlib
import os, sys
DUMP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'sgcount_b200', 'lib', 'fastx_dump')
os.chdir(__import__('tempfile').mkdtemp(prefix='sgc_fuzz_'))
def member(payload, level):
    if level < 0:  # stored
        body = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xFFFF) + payload
    else:
        c = zlib.compressobj(level, zlib.DEFLATED, -15); body = c.compress(payload) + c.flush()
    return b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25) + body + struct.pack("<II", zlib.crc32(payload), len(payload))
bad = 0
for seed in range(300):
    rng = random.Random(seed)
    fake = member(b"", 1)
    blocks = []
    for _ in range(rng.randint(1, 60)):
        n = rng.choice([0, 10, 500, 3000, 20000])
        payload = bytes(rng.choice(b"ACGT\n@+I") for _ in range(n))
        if rng.random() < 0.4:  # header look-alikes inside stored payloads
            k = rng.randint(1, 4)
            at = rng.randint(0, len(payload))
            payload = payload[:at] + fake * k + payload[at:]
        blocks.append(member(payload[:65000], rng.choice([-1, -1, 1, 6])))
    blob = b"".join(blocks)
    if rng.random() < 0.1: blob += b"garbage"
    open('f.gz', 'wb').write(blob)
    outs = []
    for t in (1, 2, 3, 5, 8, 16):
        p = subprocess.run([DUMP, 'f.gz', str(t), 'bgzfindex', '0'], capture_output=True, text=True)
        outs.append(p.stdout.split()[:2])
    if any(o != outs[0] for o in outs):
        bad += 1; print('MISMATCH seed', seed, outs)
print('done, mismatches', bad)
