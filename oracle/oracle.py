"""ctypes view of oracle/liboracle.so — the CPU parity oracle.

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The names mirror the
reference's types (Library, Permuter, Counter, Offset; /root/reference/src) so the parity
tests read like the reference's own unit tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

RC_BITTRICK = 0
RC_KEEP_N = 1

PLUS, MINUS, CENTERED, NULL = 0, 1, 2, 3

ERR_INCONSISTENT_SIZE = 1
ERR_READ_TOO_SHORT = 2
ERR_IO = 3
PANIC_DUPLICATE_SEQ = -1
PANIC_NAN = -2
PANIC_EMPTY_READER = -3
PANIC_GENEMAP = -4
PANIC_MALFORMED = -5


class OracleError(RuntimeError):
    """The reference would return Err (code > 0) or panic (code < 0)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u8p, u64, i64, sz = C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_size_t
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_records_from_memory": (C.c_int, [u8p, sz, C.c_int, C.POINTER(vp)]),
            "orc_records_from_path": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
            "orc_records_from_seqs": (C.c_int, [u8p, vp, u64, C.POINTER(vp)]),
            "orc_records_len": (u64, [vp]),
            "orc_records_seq": (C.POINTER(C.c_uint8), [vp, u64, C.POINTER(u64)]),
            "orc_records_id": (C.POINTER(C.c_uint8), [vp, u64, C.POINTER(u64)]),
            "orc_records_seq_bytes": (u64, [vp]),
            "orc_records_export_lines": (None, [vp, vp, vp]),
            "orc_records_free": (None, [vp]),
            "orc_seq_rev_comp": (None, [u8p, sz, C.c_int, u8p]),
            "orc_library_from_records": (C.c_int, [vp, C.POINTER(vp)]),
            "orc_library_len": (u64, [vp]),
            "orc_library_size": (u64, [vp]),
            "orc_library_contains": (i64, [vp, u8p, sz]),
            "orc_library_seq": (C.POINTER(C.c_uint8), [vp, u64, C.POINTER(u64)]),
            "orc_library_alias": (C.POINTER(C.c_uint8), [vp, u64, C.POINTER(u64)]),
            "orc_library_free": (None, [vp]),
            "orc_permuter_new": (C.c_int, [vp, vp, C.POINTER(vp)]),
            "orc_permuter_contains": (i64, [vp, u8p, sz]),
            "orc_permuter_map_len": (u64, [vp]),
            "orc_permuter_null_len": (u64, [vp]),
            "orc_permuter_null_contains": (C.c_int, [vp, u8p, sz]),
            "orc_permuter_free": (None, [vp]),
            "orc_bounds": (C.c_int, [u64, u64, u64, C.c_int, C.POINTER(u64), C.POINTER(u64)]),
            "orc_assign": (i64, [vp, vp, u8p, sz, C.c_int, u64, C.c_int, C.c_int]),
            "orc_counter_new": (C.c_int, [vp, vp, vp, C.c_int, u64, C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]),
            "orc_counter_get_value": (u64, [vp, u8p, sz]),
            "orc_counter_total_reads": (u64, [vp]),
            "orc_counter_matched_reads": (u64, [vp]),
            "orc_counter_counts_by_index": (None, [vp, vp, vp]),
            "orc_counter_free": (None, [vp]),
            "orc_position_counts": (C.c_int, [vp, u64, vp, C.POINTER(u64)]),
            "orc_positional_entropy": (C.c_int, [vp, u64, vp, C.POINTER(u64)]),
            "orc_entropy_from_counts": (C.c_int, [vp, u64, vp]),
            "orc_minimize_mse": (C.c_int, [vp, u64, vp, u64, C.POINTER(C.c_int), C.POINTER(u64)]),
            "orc_entropy_offset": (C.c_int, [vp, vp, u64, C.POINTER(C.c_int), C.POINTER(u64)]),
            "orc_render_results": (C.c_int, [vp, vp, u64, vp, u8p, sz, C.c_int, C.POINTER(vp)]),
            "orc_generate_sample_names": (C.c_int, [vp, u64, C.POINTER(vp)]),
            "orc_free": (None, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def _buf(b: bytes):
    return C.cast(C.c_char_p(b), C.c_void_p)


@dataclass(frozen=True)
class Offset:
    """offsetter.rs:10-34"""

    reverse: bool
    index: int

    @staticmethod
    def Forward(i: int) -> "Offset":
        return Offset(False, i)

    @staticmethod
    def Reverse(i: int) -> "Offset":
        return Offset(True, i)

    def is_forward(self) -> bool:
        return not self.reverse

    def is_reverse(self) -> bool:
        return self.reverse

    def __repr__(self):
        return f"{'Reverse' if self.reverse else 'Forward'}({self.index})"


class Records:
    """What the reference receives from fxread::initialize_reader."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_records_free(self._h)
            self._h = None

    @staticmethod
    def from_bytes(data: bytes, gz: bool = False) -> "Records":
        h = C.c_void_p()
        _check(lib().orc_records_from_memory(_buf(data), len(data), int(gz), C.byref(h)))
        return Records(h)

    @staticmethod
    def from_path(path: str) -> "Records":
        h = C.c_void_p()
        _check(lib().orc_records_from_path(path.encode(), C.byref(h)))
        return Records(h)

    @staticmethod
    def from_seqs(seqs: Sequence[bytes]) -> "Records":
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        return Records.from_packed(np.frombuffer(b"".join(seqs), dtype=np.uint8), off)

    @staticmethod
    def from_packed(seqs: np.ndarray, off: np.ndarray) -> "Records":
        """seqs: uint8 concatenation (no separators); off: uint64[n+1]."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        h = C.c_void_p()
        _check(lib().orc_records_from_seqs(seqs.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(h)))
        return Records(h)

    @staticmethod
    def from_lines(lines: np.ndarray, off: np.ndarray) -> "Records":
        """lines: newline-terminated sequence lines; off: uint64[n+1] line starts (+ end)."""
        n = len(off) - 1
        lens = (off[1:] - off[:-1] - 1).astype(np.uint64)
        mask = np.ones(len(lines), dtype=bool)
        mask[(off[1:] - 1).astype(np.int64)] = False
        packed = np.ascontiguousarray(lines[mask])
        poff = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(lens, out=poff[1:])
        return Records.from_packed(packed, poff)

    def __len__(self):
        return int(lib().orc_records_len(self._h))

    def seq(self, i: int) -> bytes:
        n = C.c_uint64()
        p = lib().orc_records_seq(self._h, i, C.byref(n))
        return C.string_at(p, n.value)

    def id(self, i: int) -> bytes:
        n = C.c_uint64()
        p = lib().orc_records_id(self._h, i, C.byref(n))
        return C.string_at(p, n.value)

    def lines(self):
        """(uint8 newline-terminated lines, uint64[n+1] offsets) — the CUDA path's input format."""
        n = len(self)
        total = int(lib().orc_records_seq_bytes(self._h)) + n
        buf = np.empty(total, dtype=np.uint8)
        off = np.empty(n + 1, dtype=np.uint64)
        lib().orc_records_export_lines(self._h, buf.ctypes.data, off.ctypes.data)
        return buf, off


def seq_rev_comp(seq: bytes, rc_mode: int = RC_BITTRICK) -> bytes:
    out = C.create_string_buffer(len(seq))
    lib().orc_seq_rev_comp(_buf(seq), len(seq), rc_mode, C.cast(out, C.c_void_p))
    return out.raw


class Library:
    """library.rs:9-99"""

    def __init__(self, handle, records):
        self._h = handle
        self._records = records

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_library_free(self._h)
            self._h = None

    @staticmethod
    def from_reader(records: Records) -> "Library":
        h = C.c_void_p()
        _check(lib().orc_library_from_records(records._h, C.byref(h)))
        return Library(h, records)

    def __len__(self):
        return int(lib().orc_library_len(self._h))

    def size(self) -> int:
        return int(lib().orc_library_size(self._h))

    def contains_index(self, token: bytes) -> int:
        return int(lib().orc_library_contains(self._h, _buf(token), len(token)))

    def contains(self, token: bytes) -> Optional[bytes]:
        i = self.contains_index(token)
        return None if i < 0 else self.alias_at(i)

    def seq_at(self, i: int) -> bytes:
        n = C.c_uint64()
        p = lib().orc_library_seq(self._h, i, C.byref(n))
        return C.string_at(p, n.value)

    def alias_at(self, i: int) -> bytes:
        n = C.c_uint64()
        p = lib().orc_library_alias(self._h, i, C.byref(n))
        return C.string_at(p, n.value)

    def keys(self):
        return [self.seq_at(i) for i in range(len(self))]

    def values(self):
        return [self.alias_at(i) for i in range(len(self))]


class Permuter:
    """permutes.rs:34-158 (literal insert algorithm)."""

    def __init__(self, handle, library):
        self._h = handle
        self._library = library

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_permuter_free(self._h)
            self._h = None

    @staticmethod
    def new(library: Library, order: Optional[Iterable[int]] = None) -> "Permuter":
        h = C.c_void_p()
        arr = None
        if order is not None:
            arr = np.ascontiguousarray(list(order), dtype=np.uint64)
            assert len(arr) == len(library)
        _check(lib().orc_permuter_new(library._h, None if arr is None else arr.ctypes.data, C.byref(h)))
        return Permuter(h, library)

    def contains_index(self, token: bytes) -> int:
        return int(lib().orc_permuter_contains(self._h, _buf(token), len(token)))

    def contains(self, token: bytes) -> Optional[bytes]:
        i = self.contains_index(token)
        return None if i < 0 else self._library.seq_at(i)

    def map_len(self) -> int:
        return int(lib().orc_permuter_map_len(self._h))

    def null_len(self) -> int:
        return int(lib().orc_permuter_null_len(self._h))

    def null_contains(self, token: bytes) -> bool:
        return bool(lib().orc_permuter_null_contains(self._h, _buf(token), len(token)))


def bounds(seq_len: int, offset: int, size: int, position: int):
    """counter.rs:158-180 -> (min, max) or None"""
    a, b = C.c_uint64(), C.c_uint64()
    ok = lib().orc_bounds(seq_len, offset, size, position, C.byref(a), C.byref(b))
    return (a.value, b.value) if ok else None


def assign(library: Library, permuter: Optional[Permuter], read: bytes, offset: Offset,
           position_recursion: bool = True, rc_mode: int = RC_BITTRICK) -> int:
    """counter.rs:96-140 for one read -> library index or -1"""
    return int(lib().orc_assign(library._h, permuter._h if permuter else None, _buf(read), len(read),
                                int(offset.reverse), offset.index, int(position_recursion), rc_mode))


class Counter:
    """counter.rs:17-252"""

    def __init__(self, handle, library, assignments):
        self._h = handle
        self._library = library
        self.assignments = assignments

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_counter_free(self._h)
            self._h = None

    @staticmethod
    def new(reader: Records, library: Library, permuter: Optional[Permuter], offset: Offset,
            size: Optional[int] = None, position_recursion: bool = True, *, rc_mode: int = RC_BITTRICK,
            n_threads: int = 1, want_assignments: bool = False) -> "Counter":
        assert size is None or size == library.size()  # count.rs:31 always passes library.size()
        h = C.c_void_p()
        assign_out = np.empty(len(reader), dtype=np.int32) if want_assignments else None
        _check(lib().orc_counter_new(reader._h, library._h, permuter._h if permuter else None,
                                     int(offset.reverse), offset.index, int(position_recursion), rc_mode,
                                     n_threads, None if assign_out is None else assign_out.ctypes.data,
                                     C.byref(h)))
        return Counter(h, library, assign_out)

    def get_value(self, alias: bytes) -> int:
        return int(lib().orc_counter_get_value(self._h, _buf(alias), len(alias)))

    def total_reads(self) -> int:
        return int(lib().orc_counter_total_reads(self._h))

    def matched_reads(self) -> int:
        return int(lib().orc_counter_matched_reads(self._h))

    def fraction_mapped(self) -> float:
        return self.matched_reads() / self.total_reads()

    def counts_by_index(self) -> np.ndarray:
        out = np.zeros(len(self._library), dtype=np.uint64)
        lib().orc_counter_counts_by_index(self._h, self._library._h, out.ctypes.data)
        return out


def position_counts(reader: Records, take: int = 2**63) -> np.ndarray:
    size = C.c_uint64()
    _check(lib().orc_position_counts(reader._h, take, None, C.byref(size)))
    out = np.zeros((size.value, 4), dtype=np.float64)
    _check(lib().orc_position_counts(reader._h, take, out.ctypes.data, C.byref(size)))
    return out


def entropy_from_counts(counts: np.ndarray) -> np.ndarray:
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    out = np.zeros(counts.shape[0], dtype=np.float64)
    _check(lib().orc_entropy_from_counts(counts.ctypes.data, counts.shape[0], out.ctypes.data))
    return out


def positional_entropy(reader: Records, take: int = 2**63) -> np.ndarray:
    size = C.c_uint64()
    _check(lib().orc_positional_entropy(reader._h, take, None, C.byref(size)))
    out = np.zeros(size.value, dtype=np.float64)
    _check(lib().orc_positional_entropy(reader._h, take, out.ctypes.data, C.byref(size)))
    return out


def minimize_mse(reference: np.ndarray, comparison: np.ndarray) -> Offset:
    reference = np.ascontiguousarray(reference, dtype=np.float64)
    comparison = np.ascontiguousarray(comparison, dtype=np.float64)
    rev, idx = C.c_int(), C.c_uint64()
    _check(lib().orc_minimize_mse(reference.ctypes.data, len(reference), comparison.ctypes.data,
                                  len(comparison), C.byref(rev), C.byref(idx)))
    return Offset(bool(rev.value), int(idx.value))


def entropy_offset(library: Records, sample: Records, subsample: int = 5000) -> Offset:
    rev, idx = C.c_int(), C.c_uint64()
    _check(lib().orc_entropy_offset(library._h, sample._h, subsample, C.byref(rev), C.byref(idx)))
    return Offset(bool(rev.value), int(idx.value))


def render_results(counters: Sequence[Counter], names: Sequence[str], library: Library,
                   genemap: Optional[bytes] = None, include_zero: bool = False) -> str:
    n = len(counters)
    hs = (C.c_void_p * n)(*[c._h for c in counters])
    nm = (C.c_char_p * n)(*[s.encode() for s in names])
    out = C.c_void_p()
    _check(lib().orc_render_results(hs, nm, n, library._h, _buf(genemap) if genemap is not None else None,
                                    len(genemap) if genemap is not None else 0, int(include_zero), C.byref(out)))
    text = C.string_at(out).decode()
    lib().orc_free(out)
    return text


def generate_sample_names(paths: Sequence[str]) -> list:
    n = len(paths)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    out = C.c_void_p()
    _check(lib().orc_generate_sample_names(arr, n, C.byref(out)))
    text = C.string_at(out).decode()
    lib().orc_free(out)
    return text.split("\n") if n else []
