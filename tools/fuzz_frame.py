import gzip, os, random, subprocess
import os, sys
DUMP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'sgcount_b200', 'lib', 'fastx_dump')
os.chdir(__import__('tempfile').mkdtemp(prefix='sgc_fuzz_'))
def fnv(seqs):
    h = 1469598103934665603
    for seq in seqs:
        for c in seq: h = ((h ^ c) * 1099511628211) & (2**64 - 1)
        h = ((h ^ 0xFF) * 1099511628211) & (2**64 - 1)
    return f"{len(seqs)} {h:x}"
bad = 0
for seed in range(400):
    rng = random.Random(seed)
    n = rng.randint(1, 400)
    lines = []
    for i in range(n):
        L = rng.choice([75, 75, 75, 20, 1, 0])
        hdr = b"@" + bytes(rng.choice(b"r0123456789 @+:/") for _ in range(rng.randint(0, 30)))
        seq = bytes(rng.choice(b"ACGTN") for _ in range(L))
        plus = b"+" if rng.random() < 0.8 else b"+" + hdr[1:]
        if rng.random() < 0.03: plus = b""                      # malformed: empty third line
        ql = L if rng.random() < 0.95 else rng.randint(0, 90)   # malformed: quality of another length
        qual = bytes(rng.choice(b"@+IF#!~") for _ in range(ql))
        lines += [hdr, seq, plus, qual]
    nl = b"\r\n" if rng.random() < 0.1 else b"\n"
    text = nl.join(lines) + (nl if rng.random() < 0.8 else b"")
    recs = text.split(b"\n")
    if recs and recs[-1] == b"": recs = recs[:-1]
    if len(recs) % 4: 
        continue  # truncated record: an error case, tested elsewhere
    want = fnv(recs[1::4])
    open('f.fq', 'wb').write(text)
    # multi-member gzip cut anywhere
    cuts = sorted(rng.sample(range(1, max(2, len(text))), min(len(text) - 1, rng.randint(0, 6)))) if len(text) > 2 else []
    with open('f.fq.gz', 'wb') as f:
        prev = 0
        for c in cuts + [len(text)]:
            f.write(gzip.compress(text[prev:c], 1)); prev = c
    for path, t in (('f.fq', 1), ('f.fq.gz', 1), ('f.fq.gz', 4)):
        p = subprocess.run([DUMP, path, str(t), 'blocks'], capture_output=True, text=True)
        if p.returncode != 0 or p.stdout.strip() != want:
            bad += 1; print('MISMATCH', seed, path, t, p.returncode, p.stdout.strip(), want, p.stderr[:200]); break
print('done, mismatches', bad)
