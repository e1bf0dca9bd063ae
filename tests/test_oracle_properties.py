"""Cross-checks inside the oracle: literal Permuter (permutes.rs:127-158) vs the closed form
the CUDA tables implement, for every key-iteration order the
reference's randomly seeded HashMap could produce, including reverse orientation, position
recursion, the early-return quirk and N handling."""
import itertools
import random

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle as orc
from oracle import oracle_py as opy


def make_library(seqs):
    recs = orc.Records.from_bytes(b"".join(b">g%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    return orc.Library.from_reader(recs)


def all_tokens(k, alphabet=b"ACGTN"):
    return [bytes(t) for t in itertools.product(alphabet, repeat=k)]


@pytest.mark.parametrize("seed", range(6))
def test_literal_permuter_equals_closed_form_for_any_order(seed):
    rng = random.Random(seed)
    k = 4
    pool = all_tokens(k, b"ACGT")
    seqs = rng.sample(pool, 12)
    library = make_library(seqs)
    for _ in range(4):
        order = list(range(len(seqs)))
        rng.shuffle(order)
        permuter = orc.Permuter.new(library, order)
        py = opy.LiteralPermuter([seqs[i] for i in order])
        assert permuter.map_len() == len(py.map) and permuter.null_len() == len(py.null)
        for tok in all_tokens(k):
            # composed lookup (library first, then permuter) — the only observable behaviour
            got = library.contains_index(tok)
            if got < 0:
                got = permuter.contains_index(tok)
            assert got == opy.closed_form_lookup(tok, seqs, True), tok


dna = st.sampled_from(b"ACGT")
noisy = st.sampled_from(b"ACGTNJacgtX")


@settings(max_examples=150, deadline=None)
@given(
    seqs=st.lists(st.lists(dna, min_size=5, max_size=5).map(bytes), min_size=1, max_size=8, unique=True),
    reads=st.lists(st.lists(noisy, min_size=0, max_size=12).map(bytes), min_size=1, max_size=12),
    reverse=st.booleans(), offset=st.integers(0, 6), recursion=st.booleans(), with_perm=st.booleans(),
    bittrick=st.booleans(),
)
def test_assign_literal_vs_closed_form(seqs, reads, reverse, offset, recursion, with_perm, bittrick):
    library = make_library(seqs)
    permuter = orc.Permuter.new(library) if with_perm else None
    rc_mode = orc.RC_BITTRICK if bittrick else orc.RC_KEEP_N
    pylib = {s: b"g%d" % i for i, s in enumerate(seqs)}
    pyperm = opy.LiteralPermuter(seqs) if with_perm else None
    off = orc.Offset(reverse, offset)
    for r in reads:
        got = orc.assign(library, permuter, r, off, recursion, rc_mode)
        want = opy.closed_form_assign(r, seqs, with_perm, reverse, offset, recursion, bittrick)
        assert got == want
        alias = opy.literal_assign(r, pylib, pyperm, reverse, offset, 5, recursion, bittrick)
        assert (alias is None) == (got < 0)
        if alias is not None:
            assert alias == b"g%d" % got


def test_early_return_quirk():
    """counter.rs:105-108: a failed trim returns None; Minus is never tried when Plus is out of range."""
    library = make_library([b"CGTAC"])
    read = b"CGTACA"[:6]  # guide sits at offset 0; with offset 1 only Minus could find it
    assert orc.assign(library, None, b"CGTACAA", orc.Offset.Forward(1), True) == 0  # len 7: C miss, P miss, M hit
    assert orc.assign(library, None, read, orc.Offset.Forward(1), True) == -1       # len 6: P out of range -> None
    assert orc.assign(library, None, read, orc.Offset.Forward(0), True) == 0


def test_reverse_counting_and_n_modes():
    library = make_library([b"AACCG"])
    permuter = orc.Permuter.new(library)
    fwd = b"TT" + b"AACCG" + b"GGG"
    rc = orc.seq_rev_comp(fwd)
    assert rc == opy.rev_comp(fwd) == b"CCCCGGTTAA"
    assert orc.assign(library, permuter, rc, orc.Offset.Reverse(2), False) == 0
    # an N inside the window of a reverse read: 'N' -> 'J' under the recalled fxread bit trick
    with_n = orc.seq_rev_comp(b"TT" + b"AANCG" + b"GGG", orc.RC_KEEP_N)
    assert orc.assign(library, permuter, with_n, orc.Offset.Reverse(2), False, orc.RC_KEEP_N) == 0
    assert orc.assign(library, permuter, with_n, orc.Offset.Reverse(2), False, orc.RC_BITTRICK) == -1
    # ... and 'J' -> 'N'
    with_j = with_n.replace(b"N", b"J")
    assert orc.assign(library, permuter, with_j, orc.Offset.Reverse(2), False, orc.RC_BITTRICK) == 0


def test_threaded_counter_equals_single_thread():
    rng = np.random.default_rng(5)
    seqs = [bytes(rng.choice(list(b"ACGT"), 6).tolist()) for _ in range(40)]
    seqs = list(dict.fromkeys(seqs))
    library = make_library(seqs)
    permuter = orc.Permuter.new(library)
    reads = []
    for _ in range(3000):
        g = bytearray(seqs[rng.integers(len(seqs))])
        if rng.random() < 0.3:
            g[rng.integers(6)] = rng.choice(list(b"ACGTN"))
        reads.append(b"GG" + bytes(g) + b"TTT")
    recs = orc.Records.from_seqs(reads)
    a = orc.Counter.new(recs, library, permuter, orc.Offset.Forward(2), want_assignments=True)
    b = orc.Counter.new(recs, library, permuter, orc.Offset.Forward(2), n_threads=4, want_assignments=True)
    assert a.total_reads() == b.total_reads() == 3000
    assert a.matched_reads() == b.matched_reads()
    assert np.array_equal(a.counts_by_index(), b.counts_by_index())
    assert np.array_equal(a.assignments, b.assignments)
    assert np.bincount(a.assignments[a.assignments >= 0], minlength=len(seqs)).tolist() == a.counts_by_index().tolist()
