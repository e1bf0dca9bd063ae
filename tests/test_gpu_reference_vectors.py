"""The reference's hot-path unit-test vectors through the CUDA C ABI.
Same cases as test_oracle_reference_vectors.py, same names as the reference's tests."""
import numpy as np
import pytest

import sgcount_b200 as sg
from sgcount_b200 import _cabi

pytestmark = pytest.mark.gpu


def batch(*seqs):
    return sg.ReadBatch.from_seqs(list(seqs), [b"seq.%d" % i for i in range(len(seqs))])


def lib_actg():
    return sg.Library.from_reader(batch(b"ACTG"))


def test_count_no_distance_no_permute():
    """counter.rs:283-288"""
    library = lib_actg()
    count = sg.Counter.new(batch(b"ACTG"), library, None, sg.Offset.Forward(0), 4, False)
    assert count.get_value(b"seq.0") == 1


def test_count_with_distance_no_permute():
    """counter.rs:291-304"""
    library = lib_actg()
    count = sg.Counter.new(batch(b"AGTG"), library, None, sg.Offset.Forward(0), 4, False)
    assert count.get_value(b"seq.0") == 0
    assert count.total_reads() == 1 and count.matched_reads() == 0


def test_count_with_distance_with_permute():
    """counter.rs:307-320"""
    library = lib_actg()
    count = sg.Counter.new(batch(b"AGTG"), library, sg.Permuter.new(library), sg.Offset.Forward(0), 4, False)
    assert count.get_value(b"seq.0") == 1
    assert count.total_reads() == 1 and count.matched_reads() == 1


def test_bounds_through_the_kernel():
    """counter.rs:323-382: (min,max) per position and None when max > len, observed through
    which window a one-guide library matches"""
    library = sg.Library.from_reader(batch(b"CATT"))
    def hit(read, offset, recursion=True):
        c = sg.Counter.new(batch(read), library, None, sg.Offset.Forward(offset), 4, recursion)
        return c.matched_reads()
    assert hit(b"ACTGCATTACTG", 4, False) == 1      # Null   -> (4, 8)
    assert hit(b"ACTGACATTCTG", 4) == 1             # Plus   -> (5, 9)
    assert hit(b"ACTCATTGACTG", 4) == 1             # Minus  -> (3, 7)
    assert hit(b"ACTGACATTCTG", 4, False) == 0      # no recursion: Plus is never tried
    assert hit(b"ACTGCAT", 4, False) == 0           # Null clipped: 8 > 7
    assert hit(b"ACTCATTG", 4) == 0                 # len 8: Plus (5,9) clipped -> None, Minus never tried
    assert hit(b"ACTCATTGA", 4) == 1                # len 9: Plus fits, misses, Minus hits
    assert hit(b"CATTAA", 0) == 1 and hit(b"ACATTA", 0) == 1   # offset 0: Centered / Plus
    assert hit(b"GGCATT", 3) == 0                   # 3+4 > 6


def test_permuter_validate_singleton():
    """permutes.rs:193-207"""
    library = lib_actg()
    permuter = sg.Permuter.new(library)
    truth = [b"AATG", b"ACGG", b"ACAG", b"TCTG", b"ACNG", b"NCTG", b"ACTA", b"GCTG", b"AGTG",
             b"ACTC", b"ATTG", b"ANTG", b"ACCG", b"ACTT", b"CCTG", b"ACTN"]
    idx, kind = permuter.lookup(truth)
    assert (idx == 0).all() and (kind == 2).all()
    idx, kind = permuter.lookup([b"ACTG", b"AGGG", b"NNTG", b"ACtG"])
    assert idx.tolist() == [0, -1, -1, -1] and kind.tolist() == [1, 0, 0, 0]
    info = permuter.info()
    assert info.n_variants == 12 and info.n_ambiguous == 0  # the 4 N variants need no storage


def test_permuter_validate_positive_and_negative():
    """permutes.rs:210-253"""
    library = sg.Library.from_reader(batch(b"AC", b"CG"))
    permuter = sg.Permuter.new(library)
    positives = [b"GC", b"TC", b"NC", b"AA", b"AT", b"AN", b"CA", b"CT", b"CN", b"GG", b"TG", b"NG"]
    idx, kind = permuter.lookup(positives)
    assert idx.tolist() == [0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1] and (kind == 2).all()
    idx, kind = permuter.lookup([b"AG", b"CC"])           # ambiguous -> the reference's _null
    assert idx.tolist() == [-1, -1]
    idx, kind = permuter.lookup([b"AC", b"CG"])           # parents resolve through the Library
    assert idx.tolist() == [0, 1] and kind.tolist() == [1, 1]
    info = permuter.info()
    assert info.n_variants == 8 and info.n_ambiguous == 2  # ACGT variants only


# The reference feeds `ACT, ACC, ACT` to positional_entropy as a bare reader; a Library cannot
# hold the duplicate (library.rs:92), and the first record is consumed without being counted
# (offsetter.rs:57), so a unique first record gives the same reference entropy.
READER = (b"ACT", b"ACC", b"ACT")
LIB_READER = (b"GGG", b"ACC", b"ACT")


def test_positional_counts():
    """offsetter.rs:266-283: size from the first record, which is not counted"""
    got = sg.position_counts(batch(*READER))
    assert got.tolist() == [[2, 0, 0, 0], [0, 2, 0, 0], [0, 1, 0, 1]]


def test_position_counts_with_n():
    """offsetter.rs:343-362 (true values; the reference test only pins the total)"""
    got = sg.position_counts(batch(b"ACT", b"ACC", b"ACT", b"ACN"))
    assert got.tolist() == [[3, 0, 0, 0], [0, 3, 0, 0], [1, 2, 1, 2]]
    got = sg.position_counts(batch(b"A", b"A", b"C", b"G", b"T", b"N", b"X"))
    assert got.tolist() == [[3, 3, 3, 3]]


def test_offset():
    """offsetter.rs:303-315"""
    library = sg.Library.from_reader(batch(*LIB_READER))
    assert sg.entropy_offset(library, batch(b"AACAAACT", b"AACAAACC", b"AACAAACT")) == sg.Offset.Forward(5)


def test_rc_offset():
    """offsetter.rs:318-328"""
    library = sg.Library.from_reader(batch(*LIB_READER))
    assert sg.entropy_offset(library, batch(b"AGTTTGTT", b"GGTTTGTT", b"AGTTTGTT")) == sg.Offset.Reverse(5)


def test_undersized_reads():
    """offsetter.rs:259-263"""
    library = sg.Library.from_reader(batch(b"ACTGACTG", b"ACTGACTT", b"ACTGACTA"))
    with pytest.raises(sg.SgcError) as e:
        sg.entropy_offset(library, batch(b"ACT", b"ACC", b"ACT"))
    assert e.value.code == _cabi.ERR_READ_TOO_SHORT


def test_nan_entropy_is_an_error():
    """offsetter.rs:123-141: a position nobody covers -> 0/0 -> NaN -> panic"""
    library = sg.Library.from_reader(batch(*LIB_READER))
    with pytest.raises(sg.SgcError) as e:
        sg.entropy_offset(library, sg.ReadBatch.from_seqs([b"AACAAACT", b"AACAA", b"AACAAAC"]))
    assert e.value.code == _cabi.ERR_NAN_ENTROPY


def test_library():
    """library.rs:119-136"""
    library = lib_actg()
    assert library.size() == 4 and len(library) == 1
    assert library.contains(b"ACTG") == b"seq.0"
    assert library.contains(b"ACTT") is None
    with pytest.raises(sg.SgcError) as e:
        sg.Library.from_reader(batch(b"ACTG", b"ACTG"))
    assert e.value.code == _cabi.ERR_DUPLICATE_SEQUENCE
    with pytest.raises(sg.SgcError):
        sg.Library.from_reader(sg.ReadBatch.from_seqs([b"ACTG", b"ACT"]))


def test_library_accepts_what_two_bits_cannot_hold():
    """library.rs:65-99 keeps opaque byte strings of any length: an N in a guide, or 31 bases, is a
    legal library (byte-keyed index, opaque.cu), and behaves like any other"""
    library = sg.Library.from_reader(batch(b"ACTG", b"ACNG"))
    assert library.contains(b"ACNG") == b"seq.1" and library.contains(b"ACTG") == b"seq.0"
    assert library.contains(b"ACAG") is None
    permuter = sg.Permuter.new(library)
    assert permuter.contains(b"ACAG") is None          # parents ACTG and ACNG: the Permuter's null set
    assert permuter.contains(b"CCNG") == b"ACNG"
    assert permuter.contains(b"ACNN") == b"ACNG"       # N is in the lexicon (permutes.rs:3)
    assert permuter.contains(b"ACNg") is None          # g is not
    long = sg.Library.from_reader(batch(b"A" * 31, b"C" * 31))
    assert long.contains(b"A" * 31) == b"seq.0" and long._exact.info().opaque == 1
    assert sg.Permuter.new(long).contains(b"A" * 30 + b"G") == b"A" * 31
