"""CUDA path vs the oracle on the same inputs, through the C ABI.

Bit-exact bar: per-read assignment, per-guide counts, total/matched, detected Offset."""
import hashlib
import os

import numpy as np
import pytest

import sgcount_b200 as sg
from oracle import oracle as orc
from sgcount_b200 import _cabi

from helpers import make_library, make_reads, oracle_library

pytestmark = pytest.mark.gpu

ACGT_BYTES = np.frombuffer(b"ACGT", dtype=np.uint8)
FIXTURES = ["sequence", "zero.sequence", "diff.sequence", "offset", "offset_clipped"]


def torch_dev(arr):
    import torch

    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def gpu_assign(library, permuter, batch, offset, recursion=True, rc_mode=_cabi.RC_BITTRICK, pad=0):
    """per-read assignment + (counts, total, matched) from sgc_counter_submit_device"""
    import torch

    lines = batch.lines
    if pad:
        lines = np.concatenate([lines, np.zeros(pad, np.uint8)])
    d_lines = torch_dev(lines)
    d_off = None if batch.line_off is None else torch_dev(batch.line_off)
    d_assign = torch.full((max(len(batch), 1),), -7, dtype=torch.int32, device="cuda")
    c = sg.Counter(library, permuter, offset, recursion, rc_mode)
    c.submit_device(d_lines.data_ptr(), lines.nbytes, len(batch), batch.stride, batch.read_len,
                    None if d_off is None else d_off.data_ptr(), d_assign.data_ptr())
    counts, total, matched = c.finish()
    return d_assign.cpu().numpy()[:len(batch)], counts, total, matched, c.launch_info()


def oracle_count(guides, seqs, with_perm, offset, recursion=True, rc_mode=orc.RC_BITTRICK, n_threads=4):
    olib, _ = oracle_library(guides)
    operm = orc.Permuter.new(olib) if with_perm else None
    oc = orc.Counter.new(orc.Records.from_seqs(seqs), olib, operm, orc.Offset(offset.reverse, offset.index),
                         None, recursion, rc_mode=rc_mode, n_threads=n_threads, want_assignments=True)
    return oc.assignments, oc.counts_by_index(), oc.total_reads(), oc.matched_reads()


# ---- config 1: the example fixtures ---------------------------------------------------------

@pytest.fixture(scope="module")
def example_library(example_dir):
    reader = sg.read_fastx(os.path.join(example_dir, "library.fasta.gz"))
    library = sg.Library.from_reader(reader)
    return library, sg.Permuter.new(library)


@pytest.mark.parametrize("name", FIXTURES)
def test_example_offset_detection(name, example_dir, example_library):
    library, _ = example_library
    reads = sg.read_fastx(os.path.join(example_dir, name + ".fastq.gz"))
    assert sg.entropy_offset(library, reads, 5000) == sg.Offset.Forward(5)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("name", FIXTURES)
def test_example_counts(name, exact, example_dir, example_library, expected):
    library, permuter = example_library
    fx = expected["fixtures"][name]
    reads = sg.read_fastx(os.path.join(example_dir, name + ".fastq.gz"))
    counter = sg.Counter.new(reads, library, None if exact else permuter, sg.Offset.Forward(5), 20, True)
    assert counter.total_reads() == fx["total_reads"]
    assert counter.matched_reads() == fx["matched_reads"]
    assert counter.counts_by_index().tolist() == fx["counts"]
    rows = "\n".join(f"{a.decode()}\t{c}" for a, c in sorted(zip(library.values(), counter.counts_by_index())))
    assert hashlib.sha256(rows.encode()).hexdigest()[:16] == fx["sha256_16"]
    assert counter.get_value(b"lib.0") == fx["counts"][0]


def test_example_permuter_stats(example_library):
    _, permuter = example_library
    info = permuter.info()
    assert info.n_variants == 100 * 20 * 3 and info.n_ambiguous == 0


# ---- synthetic: every class of read, both kernels, both orientations ------------------------

CASES = [
    # k, n_guides, read_len, offset, reverse, variable
    (20, 300, 75, 5, False, False),
    (20, 300, 75, 5, True, False),
    (20, 300, 80, 0, False, False),     # stride 81: unaligned spans in the staged kernel
    (20, 300, 80, 23, True, False),
    (20, 300, 75, 12, False, True),     # variable length -> generic kernel
    (20, 300, 75, 12, True, True),
    (5, 40, 30, 3, False, False),       # dense library: many ambiguous variants
    (5, 40, 30, 3, True, True),
    (12, 200, 50, 37, False, False),    # window flush with the read end: Plus never fits
    (21, 200, 75, 5, False, False),     # wide slots
    (25, 200, 75, 7, True, False),
    (30, 100, 64, 1, False, True),
    (30, 100, 64, 33, True, False),
    (16, 300, 75, 9, False, False),     # last k of the 4-word kernel, balanced seed parts
    (17, 300, 75, 9, True, False),      # first k of the fixed seed parts: a third part of one base
    (24, 200, 75, 2, False, False),     # last k of the 6-word wide kernel
    (24, 200, 75, 40, True, False),
]


@pytest.mark.parametrize("with_perm", [True, False])
@pytest.mark.parametrize("k,n_guides,read_len,offset,reverse,variable", CASES)
def test_per_read_assignment_matches_oracle(k, n_guides, read_len, offset, reverse, variable, with_perm):
    rng = np.random.default_rng(k * 1000 + offset + 7 * reverse + 3 * variable)
    guides = make_library(rng, n_guides, k, plant=0.05)
    wild = b"J" if reverse else b"N"
    seqs = make_reads(rng, guides, 3000, read_len, offset, reverse, variable, wild=wild)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library) if with_perm else None
    off = sg.Offset(reverse, offset)
    for recursion in (True, False):
        batch = sg.ReadBatch.from_seqs(seqs)
        got = gpu_assign(library, permuter, batch, off, recursion)
        want = oracle_count(guides, seqs, with_perm, off, recursion)
        assert np.array_equal(got[0], want[0])
        assert np.array_equal(got[1], want[1])
        assert got[2:4] == want[2:4]
        assert got[3] == int(got[1].sum())
        if not variable:
            assert got[4].kernel == 0  # the staged kernel took the whole tiles
            # the same reads as a variable-length batch go through the generic kernel
            alt = gpu_assign(library, permuter, sg.ReadBatch.from_seqs(seqs, force_offsets=True), off, recursion)
            assert alt[4].kernel == 1
            assert np.array_equal(alt[0], got[0]) and np.array_equal(alt[1], got[1])


@pytest.mark.parametrize("rc_mode", [_cabi.RC_BITTRICK, _cabi.RC_KEEP_N])
def test_reverse_complement_n_modes(rc_mode):
    """under the fxread bit trick N -> J and J -> N on reverse reads"""
    rng = np.random.default_rng(11)
    guides = make_library(rng, 100, 20)
    seqs = make_reads(rng, guides, 2000, 75, 9, reverse=True, wild=b"N")
    seqs += make_reads(rng, guides, 2000, 75, 9, reverse=True, wild=b"J")
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    off = sg.Offset.Reverse(9)
    got = gpu_assign(library, permuter, sg.ReadBatch.from_seqs(seqs), off, True, rc_mode)
    want = oracle_count(guides, seqs, True, off, True, rc_mode)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_host_submit_equals_device_submit_and_accumulates():
    rng = np.random.default_rng(3)
    guides = make_library(rng, 500, 20)
    seqs = make_reads(rng, guides, 5000, 75, 5)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    want = oracle_count(guides, seqs, True, sg.Offset.Forward(5))
    for force in (False, True):
        batch = sg.ReadBatch.from_seqs(seqs, force_offsets=force)
        c = sg.Counter(library, permuter, sg.Offset.Forward(5))
        c.submit(batch)
        c.submit(batch)  # a second batch accumulates, like a longer file would
        counts, total, matched = c.finish()
        assert np.array_equal(counts, 2 * want[1]) and total == 2 * want[2] and matched == 2 * want[3]
        c.reset()
        c.submit(batch)
        counts, total, matched = c.finish()
        assert np.array_equal(counts, want[1]) and total == want[2] and matched == want[3]


def test_empty_and_tiny_batches():
    library = sg.Library([b"ACGTACGTACGTACGTACGT"], [b"g0"])
    c = sg.Counter(library, None, sg.Offset.Forward(0))
    c.submit(sg.ReadBatch(np.zeros(0, np.uint8), 0, None, 21, 20))
    assert c.finish()[1:] == (0, 0)
    c.submit(sg.ReadBatch.from_seqs([b""], force_offsets=True))
    c.submit(sg.ReadBatch.from_seqs([b"ACGTACGTACGTACGTACGT"]))
    counts, total, matched = c.finish()
    assert counts.tolist() == [1] and total == 2 and matched == 1


def test_unaligned_device_buffer_falls_back_to_generic_kernel():
    import torch

    rng = np.random.default_rng(5)
    guides = make_library(rng, 200, 20)
    seqs = make_reads(rng, guides, 2048, 75, 5)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    batch = sg.ReadBatch.from_seqs(seqs)
    d = torch.zeros(batch.lines.nbytes + 16, dtype=torch.uint8, device="cuda")
    d[3:3 + batch.lines.nbytes] = torch.from_numpy(batch.lines).cuda()
    c = sg.Counter(library, None, sg.Offset.Forward(5))
    c.submit_device(d.data_ptr() + 3, batch.lines.nbytes, len(batch), batch.stride, batch.read_len)
    got = c.finish()
    assert c.launch_info().kernel == 1
    want = oracle_count(guides, seqs, False, sg.Offset.Forward(5))
    assert np.array_equal(got[0], want[1])


def test_brunello_shaped_two_million_reads():
    """config-2 shape (77 441 x 20 bp, Forward(5), one-mismatch on) at 2 M reads: the size the
    oracle still finishes in seconds on a few threads"""
    rng = np.random.default_rng(0xB2000002)
    guides = make_library(rng, 77441, 20, plant=0.005)
    library = sg.Library(guides, [b"lib.%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    info = permuter.info()
    assert info.n_variants + 2 * info.n_ambiguous <= 77441 * 60
    base = make_reads(rng, guides, 20000, 75, 5)
    seqs = [base[i] for i in rng.integers(0, len(base), 2_000_000)]
    off = sg.Offset.Forward(5)
    got = gpu_assign(library, permuter, sg.ReadBatch.from_seqs(seqs), off)
    want = oracle_count(guides, seqs, True, off, n_threads=os.cpu_count() or 4)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1]) and got[2:4] == want[2:4]


def test_count_replicas_give_the_same_table_on_a_skewed_sample():
    """sgc_counter_set_replicas: a third of the reads carry one guide; 1, 16 and 64 copies of the
    count vector must fold to the same table, batch after batch"""
    rng = np.random.default_rng(21)
    guides = make_library(rng, 400, 20)
    seqs = make_reads(rng, guides, 6000, 75, 5)
    hot = seqs[0][:5] + guides[7] + seqs[0][25:]
    seqs = [hot if rng.random() < 0.33 else s for s in seqs]
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    want = oracle_count(guides, seqs, True, sg.Offset.Forward(5))
    for force in (False, True):
        batch = sg.ReadBatch.from_seqs(seqs, force_offsets=force)
        for replicas in (1, 16, 64, 3):
            c = sg.Counter(library, permuter, sg.Offset.Forward(5))
            c.set_replicas(replicas)
            c.submit(batch)
            c.submit(batch)
            counts, total, matched = c.finish()
            assert np.array_equal(counts, 2 * want[1]) and (total, matched) == (2 * want[2], 2 * want[3])
            c.set_replicas(1)
            c.submit(batch)
            assert np.array_equal(c.finish()[0], 3 * want[1])


def test_randomised_geometries_match_oracle():
    """Forty random (k, library size and density, read length, offset, orientation, fixed/variable
    length, Permuter, recursion, rc_mode) combinations, 1 500 reads each, per-read assignment
    against the oracle: the kernel families (4/5/6/8 window words), both seed-part policies,
    windows flush with either end of the read, odd strides, dense libraries full of ambiguous
    variants."""
    rng = np.random.default_rng(20250711)
    for case in range(40):
        k = int(rng.choice([4, 7, 11, 15, 16, 17, 18, 19, 20, 20, 20, 21, 23, 24, 25, 28, 30]))
        n_guides = int(rng.choice([30, 200, 1500])) if k > 6 else int(rng.choice([20, 60]))
        read_len = int(rng.integers(k + 1, k + 70))
        offset = int(rng.integers(0, read_len - k + 1))
        reverse, variable = bool(rng.integers(2)), bool(rng.random() < 0.3)
        with_perm, recursion = bool(rng.random() < 0.75), bool(rng.random() < 0.75)
        rc_mode = _cabi.RC_BITTRICK if rng.random() < 0.7 else _cabi.RC_KEEP_N
        guides = make_library(rng, n_guides, k, plant=float(rng.choice([0.0, 0.02, 0.15])))
        wild = b"J" if (reverse and rc_mode == _cabi.RC_BITTRICK) else b"N"
        seqs = make_reads(rng, guides, 1500, read_len, offset, reverse, variable, wild=wild)
        library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
        permuter = sg.Permuter.new(library) if with_perm else None
        off = sg.Offset(reverse, offset)
        got = gpu_assign(library, permuter, sg.ReadBatch.from_seqs(seqs), off, recursion, rc_mode)
        want = oracle_count(guides, seqs, with_perm, off, recursion, rc_mode)
        label = (case, k, n_guides, read_len, offset, reverse, variable, with_perm, recursion, rc_mode)
        assert np.array_equal(got[0], want[0]), label
        assert np.array_equal(got[1], want[1]) and got[2:4] == want[2:4], label
        assert got[4].kernel == (1 if variable else 0), label


# ---- skew plan, launch splitting, read shards ------------------------------------------------

def _skewed_sample(rng, guides, n_reads, share, offset=5, read_len=75):
    """n_reads reads drawn from 20 000 distinct ones, `share` of them carrying guide 7 exactly"""
    base = make_reads(rng, guides, 20000, read_len, offset)
    hot = base[0][:offset] + guides[7] + base[0][offset + len(guides[7]):]
    pick = rng.integers(0, len(base), n_reads)
    is_hot = rng.random(n_reads) < share
    return [hot if h else base[i] for h, i in zip(is_hot, pick)]


@pytest.mark.parametrize("share", [0.0, 0.1, 0.9])
def test_skew_plan_is_automatic_and_does_not_change_the_table(share):
    """counter.rs:232-235 costs the same whatever the abundances; here a counter samples its first
    batch and, when a guide stands out, spreads the atomics (16 replicas) and keeps the guides
    with >= 1 % of the reads in registers.  No caller action; same table as the oracle."""
    import torch

    rng = np.random.default_rng(77)
    guides = make_library(rng, 3000, 20)
    seqs = _skewed_sample(rng, guides, 600_000, share)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    want = oracle_count(guides, seqs, True, sg.Offset.Forward(5), n_threads=os.cpu_count() or 4)
    batch = sg.ReadBatch.from_seqs(seqs)
    d = torch_dev(np.concatenate([batch.lines, np.zeros(64, np.uint8)]))
    c = sg.Counter(library, permuter, sg.Offset.Forward(5))
    for rounds in (1, 2):
        c.submit_device(d.data_ptr(), batch.lines.nbytes, len(batch), batch.stride, batch.read_len)
        counts, total, matched = c.finish()
        assert np.array_equal(counts, rounds * want[1]) and (total, matched) == (rounds * want[2], rounds * want[3])
    info = c.launch_info()
    if share > 0.0:  # (with 3 000 log-normal guides the top one may pass 1/256 of the reads by itself)
        assert info.replicas == 16 and info.hot_guides >= 1
    # the host path plans on its first chunk as well
    h = sg.Counter(library, permuter, sg.Offset.Forward(5))
    h.submit(batch)
    assert np.array_equal(h.finish()[0], want[1])
    assert h.launch_info().replicas == info.replicas
    # and the plan can be switched off
    off = sg.Counter(library, permuter, sg.Offset.Forward(5))
    off.set_replicas(1)
    off.submit_device(d.data_ptr(), batch.lines.nbytes, len(batch), batch.stride, batch.read_len)
    assert np.array_equal(off.finish()[0], want[1]) and off.launch_info().replicas == 1


def test_a_batch_is_cut_into_several_launches(monkeypatch):
    """count.cu cuts a batch into launches of at most 2^30 - 2 M reads (the parked-read queue
    keeps a 30-bit read index); SGC_MAX_LAUNCH_TILES lowers the cap so the splitting runs here:
    57 tiles of 32 reads per launch over 50 021 reads = 28 streaming launches + the remainder."""
    rng = np.random.default_rng(8)
    guides = make_library(rng, 800, 20)
    seqs = make_reads(rng, guides, 50_021, 75, 5)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    want = oracle_count(guides, seqs, True, sg.Offset.Forward(5))
    monkeypatch.setenv("SGC_MAX_LAUNCH_TILES", "57")
    batch = sg.ReadBatch.from_seqs(seqs)
    got = gpu_assign(library, permuter, batch, sg.Offset.Forward(5), pad=64)
    assert got[4].launches_total == -(-(50_021 // 32) // 57) + 1
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2:4] == want[2:4]
    c = sg.Counter(library, permuter, sg.Offset.Forward(5))  # production mode, host path
    c.submit(batch)
    assert np.array_equal(c.finish()[0], want[1])


def test_reduce_counts_of_read_shards_on_one_device():
    """sgc_reduce_counts: three read shards of one sample, three counters, summed into shard 1"""
    rng = np.random.default_rng(12)
    guides = make_library(rng, 600, 20)
    seqs = make_reads(rng, guides, 9000, 75, 5)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    want = oracle_count(guides, seqs, True, sg.Offset.Forward(5))
    shards = []
    for part in (seqs[:2000], seqs[2000:2001], seqs[2001:]):
        c = sg.Counter(library, permuter, sg.Offset.Forward(5))
        c.submit(sg.ReadBatch.from_seqs(part))
        shards.append(c)
    sg.reduce_counts(shards, root=1)
    counts, total, matched = shards[1].finish()
    assert np.array_equal(counts, want[1]) and (total, matched) == (want[2], want[3])
    with pytest.raises(sg.SgcError):
        sg.reduce_counts([shards[0], shards[0]])


@pytest.mark.parametrize("path", ["peer copies", "nccl"])  # in this order: sgc_reduce_prepare switches the process to NCCL
def test_reduce_counts_across_devices(path):
    """read shards on every device of the box summed into one: peer-to-peer copies + add when nobody
    has asked for NCCL, one ncclReduce of u64[n_guides + 2] once sgc_reduce_prepare has"""
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two devices")
    if path == "nccl":
        sg.reduce_prepare(list(range(n_dev)))
    rng = np.random.default_rng(13)
    guides = make_library(rng, 2000, 20)
    seqs = make_reads(rng, guides, 40_000, 75, 5)
    want = oracle_count(guides, seqs, True, sg.Offset.Forward(5))
    cuts = np.linspace(0, len(seqs), 2 * n_dev + 1).astype(int)  # two shards per device: fold + NCCL
    shards, keep = [], []
    for i in range(2 * n_dev):
        dev = i % n_dev
        library = sg.Library(guides, [b"g%d" % j for j in range(len(guides))], device=dev)
        permuter = sg.Permuter.new(library)
        c = sg.Counter(library, permuter, sg.Offset.Forward(5))
        c.submit(sg.ReadBatch.from_seqs(seqs[cuts[i]:cuts[i + 1]]))
        shards.append(c)
        keep.append((library, permuter))
    for root in (0, 2 * n_dev - 1):
        sg.reduce_counts(shards, root=root)
        counts, total, matched = shards[root].finish()
        assert np.array_equal(counts, want[1]) and (total, matched) == (want[2], want[3])
        if root == 0:  # count everything again for the second root
            for i, c in enumerate(shards):
                c.reset()
                c.submit(sg.ReadBatch.from_seqs(seqs[cuts[i]:cuts[i + 1]]))


# ---- offset detector: ties ---------------------------------------------------------------------

def test_offset_tie_goes_to_reverse_and_near_ties_follow_the_oracle():
    """assign_offset (offsetter.rs:143-149) answers Forward only if min_f < min_r: an exact tie is
    Reverse.  Palindromic reads give a positional histogram that equals its own mirror image, so
    every forward window has a reversed twin with the same MSE bit for bit."""
    rng = np.random.default_rng(99)
    guides = make_library(rng, 300, 12)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    olib, lib_recs = oracle_library(guides)
    half = ACGT_BYTES[rng.integers(0, 4, (4000, 20))]
    # skew the composition by position so the entropy profile is not flat
    for p in range(20):
        mask = rng.random(4000) < 0.04 * p
        half[mask, p] = ord("G")
    pal = np.concatenate([half, half[:, ::-1]], axis=1)
    seqs = [r.tobytes() for r in pal]
    got = sg.entropy_offset(library, sg.ReadBatch.from_seqs(seqs), 5000)
    want = orc.entropy_offset(lib_recs, orc.Records.from_seqs(seqs), 5000)
    assert got.reverse and (got.reverse, got.index) == (want.reverse, want.index)
    # near ties: break the symmetry by a handful of bases at one position inside the winning
    # window of one direction (a one-count change decides Forward or Reverse; the oracle's answers
    # on this grid are a mix of both)
    outcomes = set()
    for flips in (1, 2, 5, 25):
        for col in (3, 8, 30, 36):
            near = pal.copy()
            rows = rng.choice(np.arange(1, 4000), flips, replace=False)
            near[rows, col] = np.where(near[rows, col] == ord("G"), ord("C"), ord("G"))
            seqs = [r.tobytes() for r in near]
            got = sg.entropy_offset(library, sg.ReadBatch.from_seqs(seqs), 5000)
            want = orc.entropy_offset(lib_recs, orc.Records.from_seqs(seqs), 5000)
            assert (got.reverse, got.index) == (want.reverse, want.index), (flips, col)
            outcomes.add(want.reverse)
    assert outcomes == {False, True}


# ---- span records --------------------------------------------------------------------------------

def test_span_records_count_like_whole_reads():
    """sgc_span_geometry: the guide window and one byte either side are all Counter::assign looks at
    (counter.rs:164-174).  For every geometry — windows flush with either end of the read, offset
    0, both orientations, recursion on and off, narrow and wide keys — span records counted under
    the span Offset give the per-read assignment of the whole reads, which is the oracle's."""
    rng = np.random.default_rng(4242)
    cases = [(20, 75, 5, False, True), (20, 75, 0, False, True), (20, 75, 55, False, True), (20, 75, 54, True, True),
             (20, 75, 0, True, True), (20, 75, 55, True, True), (20, 75, 7, True, False), (20, 20, 0, False, True),
             (16, 50, 3, False, True), (24, 60, 11, True, True), (30, 75, 44, False, True), (28, 75, 45, True, True)]
    for k, read_len, offset, reverse, recursion in cases:
        guides = make_library(rng, 300, k, plant=0.05)
        wild = b"J" if reverse else b"N"
        seqs = make_reads(rng, guides, 3000, read_len, offset, reverse, False, wild=wild)
        library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
        permuter = sg.Permuter.new(library)
        off = sg.Offset(reverse, offset)
        batch = sg.ReadBatch.from_seqs(seqs)
        want = oracle_count(guides, seqs, True, off, recursion)
        spans, span_off = sg.span_batch(batch, k, off, recursion)
        assert spans.stride % 8 == 0 and spans.read_len <= k + 2 and span_off.reverse == reverse
        got = gpu_assign(library, permuter, spans, span_off, recursion, pad=64)
        label = (k, read_len, offset, reverse, recursion, spans.read_len, span_off)
        assert np.array_equal(got[0], want[0]), label
        assert np.array_equal(got[1], want[1]) and got[2:4] == want[2:4], label
        assert got[4].kernel == 0, label  # the streaming kernel
        c = sg.Counter(library, permuter, span_off, recursion)  # production mode, host path
        c.submit(spans)
        assert np.array_equal(c.finish()[0], want[1]), label
    with pytest.raises(sg.SgcError):
        sg.span_geometry(20, 24, sg.Offset.Forward(5))  # the Centered window does not fit


# ---- libraries the 2-bit tables cannot hold -----------------------------------------------------

def _opaque_library(rng, n, k, kind):
    """kind 'long': A,C,G,T guides of k > 30 bases; 'bytes': a tenth of the guides carry an N or a
    lower-case base (library.rs keeps sequences as opaque byte strings)"""
    guides = make_library(rng, n, k, plant=0.05)
    if kind == "bytes":
        out, seen = [], set()
        for g in guides:
            if rng.random() < 0.1:
                b = bytearray(g)
                b[int(rng.integers(k))] = int(rng.choice(list(b"Nnacgt")))
                g = bytes(b)
            if g not in seen:
                seen.add(g)
                out.append(g)
        guides = out
    return guides


@pytest.mark.parametrize("kind,k", [("bytes", 20), ("bytes", 12), ("long", 31), ("long", 50), ("bytes", 40)])
def test_opaque_libraries_match_the_oracle(kind, k):
    """A library byte outside A,C,G,T or a guide longer than 30 bases takes the byte-keyed index
    (opaque.cu).  Per-read assignment, counts and the Permuter's observable lookups equal the
    oracle's in both orientations, both rc modes, fixed and variable length, recursion on and off."""
    rng = np.random.default_rng(1000 + k)
    guides = _opaque_library(rng, 400, k, kind)
    library = sg.Library(guides, [b"g%d" % i for i in range(len(guides))])
    permuter = sg.Permuter.new(library)
    assert permuter.info().opaque == 1
    olib, _ = oracle_library(guides)
    operm = orc.Permuter.new(olib)
    # composed lookup on members, their lexicon variants and noise
    tokens = []
    for g in guides[:60]:
        tokens.append(g)
        for _ in range(4):
            b = bytearray(g)
            b[int(rng.integers(k))] = int(rng.choice(list(b"ACGTNx")))
            tokens.append(bytes(b))
    idx, kinds = permuter.lookup(tokens)
    for t, i, kd in zip(tokens, idx, kinds):
        want = olib.contains_index(t)
        if want >= 0:
            assert (i, kd) == (want, 1), t
        else:
            want = operm.contains_index(t)
            assert (i, kd) == ((want, 2) if want >= 0 else (-1, 0)), t
    for case in range(6):
        read_len = int(rng.integers(k + 1, k + 50))
        offset = int(rng.integers(0, read_len - k + 1))
        reverse, variable = bool(case & 1), bool(case & 2)
        recursion, with_perm = case != 4, case != 5
        rc_mode = _cabi.RC_KEEP_N if case >= 3 else _cabi.RC_BITTRICK
        seqs = make_reads(rng, guides, 2000, read_len, offset, reverse, variable)
        off = sg.Offset(reverse, offset)
        got = gpu_assign(library, permuter if with_perm else None, sg.ReadBatch.from_seqs(seqs), off, recursion, rc_mode)
        want = oracle_count(guides, seqs, with_perm, off, recursion, rc_mode)
        label = (kind, k, case, read_len, offset, reverse, variable)
        assert np.array_equal(got[0], want[0]), label
        assert np.array_equal(got[1], want[1]) and got[2:4] == want[2:4], label
        assert got[4].kernel == 2, label
        c = sg.Counter(library, permuter if with_perm else None, off, recursion, rc_mode)
        c.submit(sg.ReadBatch.from_seqs(seqs))
        c.submit(sg.ReadBatch.from_seqs(seqs))
        assert np.array_equal(c.finish()[0], 2 * want[1]), label
    # the offset detector works on bytes and does not care how the library is indexed
    seqs = make_reads(rng, guides, 3000, k + 30, 7)
    got = sg.entropy_offset(library, sg.ReadBatch.from_seqs(seqs), 5000)
    want = orc.entropy_offset(oracle_library(guides)[1], orc.Records.from_seqs(seqs), 5000)
    assert (got.reverse, got.index) == (want.reverse, want.index)


def test_opaque_library_duplicates_and_limits():
    with pytest.raises(sg.SgcError) as e:
        sg.Library([b"ACGTNACGTA", b"TTGTNACGTA", b"ACGTNACGTA"], [b"a", b"b", b"c"])
    assert e.value.code == _cabi.ERR_DUPLICATE_SEQUENCE and "ACGTNACGTA" in str(e.value)
    with pytest.raises(sg.SgcError) as e:
        sg.Library([b"A" * 1025], [b"a"])
    assert e.value.code == _cabi.ERR_K_UNSUPPORTED
