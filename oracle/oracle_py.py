"""Pure-Python twin of the oracle, for small cases only (TEST INFRASTRUCTURE, NOT PRODUCT).

Two statements of the same path live here so they can be checked against each other and
against oracle.cpp:

* ``LiteralPermuter`` — the reference's stateful insert algorithm, permutes.rs:63-158;
* ``closed_form_assign`` — the order-independent closed form the CUDA tables implement
  (SURVEY.md appendix A): a token that is not a library member matches iff exactly one
  library sequence is at Hamming distance 1 and the token's differing byte is in ACGTN.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

LEXICON = b"ACGTN"  # permutes.rs:3


def rev_comp(seq: bytes, bittrick: bool = True) -> bytes:
    """fxread Record::seq_rev_comp as used at counter.rs:203."""
    if bittrick:
        return bytes((c ^ 4) if (c & 2) else (c ^ 21) for c in reversed(seq))
    table = {65: 84, 67: 71, 71: 67, 84: 65}
    return bytes(table.get(c, c) for c in reversed(seq))


class LiteralPermuter:
    """permutes.rs:47-158, dictionaries in place of hashbrown maps."""

    def __init__(self, sequences: Sequence[bytes]):
        self.map: Dict[bytes, bytes] = {}
        self.null = set()
        for seq in sequences:
            for idx in range(len(seq)):
                for y in LEXICON:
                    if y == seq[idx]:
                        continue
                    self._insert(seq, seq[:idx] + bytes([y]) + seq[idx + 1:])

    def _insert(self, sequence: bytes, permutation: bytes) -> None:  # permutes.rs:127-144
        if sequence not in self.null:
            self.null.add(sequence)
        if permutation not in self.null:
            if permutation in self.map:
                del self.map[permutation]
                self.null.add(permutation)
            else:
                self.map[permutation] = sequence

    def contains(self, token: bytes) -> Optional[bytes]:
        return self.map.get(token)


def bounds(seq_len: int, offset: int, size: int, position: str) -> Optional[Tuple[int, int]]:
    """counter.rs:158-180; position in {'plus','minus','centered','null'}"""
    if position == "plus":
        lo, hi = offset + 1, offset + 1 + size
    elif position == "minus":
        if offset == 0:
            return None
        lo, hi = offset - 1, offset - 1 + size
    else:
        lo, hi = offset, offset + size
    return None if hi > seq_len else (lo, hi)


def literal_assign(read: bytes, library: Dict[bytes, bytes], permuter: Optional[LiteralPermuter],
                   reverse: bool, offset: int, size: int, recursion: bool, bittrick: bool = True) -> Optional[bytes]:
    """counter.rs:96-140 -> alias or None"""
    for position in (("centered", "plus", "minus") if recursion else ("null",)):
        b = bounds(len(read), offset, size, position)
        if b is None:
            return None  # counter.rs:105-108: return, not continue
        src = rev_comp(read, bittrick) if reverse else read
        token = src[b[0]:b[1]]
        if token in library:
            return library[token]
        if permuter is not None:
            parent = permuter.contains(token)
            if parent is not None and parent in library:
                return library[parent]
    return None


def closed_form_lookup(token: bytes, lib_seqs: Sequence[bytes], with_permuter: bool) -> int:
    """library index matched by one token under the closed form, or -1"""
    for i, s in enumerate(lib_seqs):
        if s == token:
            return i
    if not with_permuter:
        return -1
    parents = []
    for i, s in enumerate(lib_seqs):
        diff = [j for j in range(len(s)) if s[j] != token[j]]
        if len(diff) == 1 and token[diff[0]] in LEXICON:
            parents.append(i)
    return parents[0] if len(parents) == 1 else -1


def closed_form_assign(read: bytes, lib_seqs: Sequence[bytes], with_permuter: bool, reverse: bool,
                       offset: int, recursion: bool, bittrick: bool = True) -> int:
    size = len(lib_seqs[0])
    for position in (("centered", "plus", "minus") if recursion else ("null",)):
        b = bounds(len(read), offset, size, position)
        if b is None:
            return -1
        src = rev_comp(read, bittrick) if reverse else read
        hit = closed_form_lookup(src[b[0]:b[1]], lib_seqs, with_permuter)
        if hit >= 0:
            return hit
    return -1


# ---- offsetter.rs:37-163 ---------------------------------------------------------------

def position_counts(seqs: Sequence[bytes]) -> List[List[float]]:
    size = len(seqs[0])  # first record consumed (offsetter.rs:57)
    m = [[0.0] * 4 for _ in range(size)]
    for s in seqs[1:]:
        for idx, c in enumerate(s[:size]):
            j = b"ACGT".find(bytes([c]))
            if j >= 0:
                m[idx][j] += 1.0
            else:
                for q in range(4):
                    m[idx][q] += 1.0
    return m


def positional_entropy(seqs: Sequence[bytes]) -> List[float]:
    out = []
    for row in position_counts(seqs):
        total = ((row[0] + row[1]) + row[2]) + row[3]
        acc = 0.0
        for v in row:
            p = v / total if total != 0 else float("nan")
            acc += 0.0 if p == 0.0 else p * math.log(p)
        out.append(-acc)
    return out


def minimize_mse(ref: Sequence[float], cmp_: Sequence[float]) -> Tuple[bool, int]:
    """-> (is_reverse, index); offsetter.rs:109-163"""
    if len(cmp_) < len(ref):
        raise ValueError("read shorter than reference")

    def windowed(b):
        out = []
        for x in range(len(b) - len(ref) + 1):
            acc = 0.0
            for i in range(len(ref)):
                d = ref[i] - b[x + i]
                acc += d * d
            out.append(acc / len(ref))
        return out

    f, r = windowed(list(cmp_)), windowed(list(reversed(cmp_)))
    if any(math.isnan(v) for v in f + r):
        raise FloatingPointError("NaN in entropy")
    mf, mr = min(f), min(r)
    return (False, f.index(mf)) if mf < mr else (True, r.index(mr))
