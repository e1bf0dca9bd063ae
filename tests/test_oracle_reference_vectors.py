"""The reference's own hot-path unit-test vectors, replayed on the oracle.

One test per pinned row; the docstring of each names the reference test it restates.
This is what pins oracle/oracle.cpp (and its Python twin) to the reference.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import oracle_py as opy


def fasta(*seqs):
    return orc.Records.from_bytes(b"".join(b">seq.%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))


def lib_actg():
    return orc.Library.from_reader(fasta(b"ACTG"))


# ---- counter.rs -------------------------------------------------------------------------

def test_count_no_distance_no_permute():
    """counter.rs:283-288"""
    library = lib_actg()
    count = orc.Counter.new(fasta(b"ACTG"), library, None, orc.Offset.Forward(0), 4, False)
    assert count.get_value(b"seq.0") == 1


def test_count_with_distance_no_permute():
    """counter.rs:291-304 (two reference tests with the same body)"""
    library = lib_actg()
    count = orc.Counter.new(fasta(b"AGTG"), library, None, orc.Offset.Forward(0), 4, False)
    assert count.get_value(b"seq.0") == 0


def test_count_with_distance_with_permute():
    """counter.rs:307-320"""
    library = lib_actg()
    permuter = orc.Permuter.new(library)
    count = orc.Counter.new(fasta(b"AGTG"), library, permuter, orc.Offset.Forward(0), 4, False)
    assert count.get_value(b"seq.0") == 1
    assert count.total_reads() == 1 and count.matched_reads() == 1


@pytest.mark.parametrize("position,want", [(orc.NULL, (4, 8)), (orc.PLUS, (5, 9)), (orc.MINUS, (3, 7))])
def test_bounds_checking(position, want):
    """counter.rs:323-352"""
    assert orc.bounds(len(b"ACTGACTGACTG"), 4, 4, position) == want


@pytest.mark.parametrize("seq,position", [(b"ACTGACT", orc.NULL), (b"ACTGACT", orc.PLUS), (b"ACTGAC", orc.MINUS)])
def test_bounds_checking_clipped(seq, position):
    """counter.rs:355-382"""
    assert orc.bounds(len(seq), 4, 4, position) is None


def test_bounds_python_twin():
    for n, pos, name in [(12, orc.NULL, "null"), (12, orc.PLUS, "plus"), (12, orc.MINUS, "minus"),
                         (7, orc.NULL, "null"), (7, orc.PLUS, "plus"), (6, orc.MINUS, "minus"),
                         (9, orc.MINUS, "minus")]:
        assert orc.bounds(n, 4, 4, pos) == opy.bounds(n, 4, 4, name)
    assert orc.bounds(10, 0, 4, orc.MINUS) is None and opy.bounds(10, 0, 4, "minus") is None


# ---- permutes.rs ------------------------------------------------------------------------

def test_permuter_validate_singleton():
    """permutes.rs:193-207"""
    library = lib_actg()
    permuter = orc.Permuter.new(library)
    truth = [b"AATG", b"ACGG", b"ACAG", b"TCTG", b"ACNG", b"NCTG", b"ACTA", b"GCTG", b"AGTG",
             b"ACTC", b"ATTG", b"ANTG", b"ACCG", b"ACTT", b"CCTG", b"ACTN"]
    assert all(permuter.contains(x) == b"ACTG" for x in truth)
    assert all(not permuter.null_contains(x) for x in truth)
    assert permuter.null_contains(b"ACTG")
    assert permuter.null_len() == 1
    assert permuter.map_len() == 16


def test_permuter_validate_positive():
    """permutes.rs:210-231"""
    permuter = orc.Permuter.new(orc.Library.from_reader(fasta(b"AC", b"CG")))
    known = [b"GC", b"TC", b"NC", b"AA", b"AT", b"AN", b"CA", b"CT", b"CN", b"GG", b"TG", b"NG"]
    assert all(permuter.contains(x) is not None for x in known)
    assert permuter.map_len() == 12
    assert all(not permuter.null_contains(x) for x in known)
    py = opy.LiteralPermuter([b"AC", b"CG"])
    assert sorted(py.map) == sorted(known)


def test_permuter_validate_negative():
    """permutes.rs:234-253"""
    permuter = orc.Permuter.new(orc.Library.from_reader(fasta(b"AC", b"CG")))
    known = [b"AG", b"CG", b"CC", b"AG"]
    assert all(permuter.null_contains(x) for x in known)
    assert permuter.null_len() == 4
    assert all(permuter.contains(x) is None for x in known)
    assert opy.LiteralPermuter([b"AC", b"CG"]).null == {b"AC", b"CG", b"CC", b"AG"}


# ---- offsetter.rs -----------------------------------------------------------------------

READER = (b"ACT", b"ACC", b"ACT")
OFFSET_READER = (b"AACAAACT", b"AACAAACC", b"AACAAACT")
RC_OFFSET_READER = (b"AGTTTGTT", b"GGTTTGTT", b"AGTTTGTT")


def test_minimization():
    """offsetter.rs:249-256"""
    got = orc.minimize_mse(np.linspace(0.0, 10.0, 11), np.linspace(10.0, 20.0, 100))
    assert got == orc.Offset.Forward(0)
    assert opy.minimize_mse(list(np.linspace(0.0, 10.0, 11)), list(np.linspace(10.0, 20.0, 100))) == (False, 0)


def test_undersized_minimization():
    """offsetter.rs:259-263"""
    with pytest.raises(orc.OracleError) as e:
        orc.minimize_mse(np.linspace(0.0, 10.0, 11), np.linspace(10.0, 20.0, 5))
    assert e.value.code == orc.ERR_READ_TOO_SHORT


def test_sequence_size_consumes_first_record():
    """offsetter.rs:266-271: size 3 from the first record, which is then not counted"""
    counts = orc.position_counts(fasta(*READER))
    assert counts.shape == (3, 4)
    assert counts.sum() == 2 * 3  # two remaining records, three positions each


def test_positional_counts():
    """offsetter.rs:274-283"""
    want = np.array([[2.0, 0, 0, 0], [0, 2.0, 0, 0], [0, 1.0, 0, 1.0]])
    assert np.array_equal(orc.position_counts(fasta(*READER)), want)
    assert opy.position_counts(READER) == want.tolist()


def test_normalize():
    """offsetter.rs:286-300 (checked through the entropy of the normalised rows)"""
    h = orc.entropy_from_counts(np.array([[2.0, 0, 0, 0], [0, 2.0, 0, 0], [0, 1.0, 0, 1.0]]))
    assert h[0] == 0.0 and h[1] == 0.0
    assert h[2] == -(0.5 * np.log(0.5) + 0.5 * np.log(0.5))


def test_offset():
    """offsetter.rs:303-315"""
    ref = orc.positional_entropy(fasta(*READER))
    cmp_ = orc.positional_entropy(fasta(*OFFSET_READER))
    assert orc.minimize_mse(ref, cmp_) == orc.Offset.Forward(5)
    assert orc.entropy_offset(fasta(*READER), fasta(*OFFSET_READER)) == orc.Offset.Forward(5)
    assert opy.minimize_mse(opy.positional_entropy(READER), opy.positional_entropy(OFFSET_READER)) == (False, 5)


def test_rc_offset():
    """offsetter.rs:318-328"""
    ref = orc.positional_entropy(fasta(*READER))
    cmp_ = orc.positional_entropy(fasta(*RC_OFFSET_READER))
    assert orc.minimize_mse(ref, cmp_) == orc.Offset.Reverse(5)
    assert opy.minimize_mse(opy.positional_entropy(READER), opy.positional_entropy(RC_OFFSET_READER)) == (True, 5)


def test_offset_enum():
    """offsetter.rs:331-340"""
    o = orc.Offset.Forward(5)
    assert o.index == 5 and o.is_forward() and not o.is_reverse()
    o = orc.Offset.Reverse(5)
    assert o.index == 5 and not o.is_forward() and o.is_reverse()


def test_base_map_and_counts_with_n():
    """offsetter.rs:343-362: A0 C1 G2 T3, anything else adds one to all four"""
    got = orc.position_counts(fasta(b"ACT", b"ACC", b"ACT", b"ACN"))
    # The reference test only asserts (posmat - expected).sum() == 0 against
    # expected = [[3,0,0,0],[0,3,0,0],[2,1,2,1]]; that holds for any matrix with the same
    # total.  Records 2-4 are ACC, ACT, ACN, so the true last row is C:2 T:2 A:1 G:1.
    ref_expected = np.array([[3.0, 0, 0, 0], [0, 3.0, 0, 0], [2.0, 1.0, 2.0, 1.0]])
    assert (got - ref_expected).sum() == 0.0
    assert np.array_equal(got, np.array([[3.0, 0, 0, 0], [0, 3.0, 0, 0], [1.0, 2.0, 1.0, 2.0]]))
    assert opy.position_counts([b"ACT", b"ACC", b"ACT", b"ACN"]) == got.tolist()
    x = orc.position_counts(fasta(b"A", b"A", b"C", b"G", b"T", b"N", b"X"))
    assert np.array_equal(x, np.array([[3.0, 3.0, 3.0, 3.0]]))


# ---- library.rs -------------------------------------------------------------------------

def test_library():
    """library.rs:119-136"""
    library = lib_actg()
    assert library.size() == 4 and len(library) == 1
    assert library.contains(b"ACTG") == b"seq.0"
    assert library.contains(b"ACTT") is None
    with pytest.raises(orc.OracleError) as e:
        orc.Library.from_reader(fasta(b"ACTG", b"ACTG"))
    assert e.value.code == orc.PANIC_DUPLICATE_SEQ


def test_library_inconsistent_sizes():
    """library.rs:79-85"""
    with pytest.raises(orc.OracleError) as e:
        orc.Library.from_reader(fasta(b"ACTG", b"ACT"))
    assert e.value.code == orc.ERR_INCONSISTENT_SIZE
