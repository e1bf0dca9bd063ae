"""ctypes view of libsgcount_synth.so — synthetic libraries and reads in the BASELINE shapes
(sgcount_b200/synth/synth.h).  Benchmark / test input generator."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsgcount_synth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: run __graft_entry__.build()")
        L = C.CDLL(_SO)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.sgs_last_error.restype = C.c_char_p
        L.sgs_make_library.argtypes = [u64, u32, u32, vp]
        L.sgs_sample_create.argtypes = [u64, u32, vp, u32, u32, u32, u32, C.c_int, C.POINTER(vp)]
        L.sgs_sample_destroy.argtypes = [vp]
        L.sgs_sample_destroy.restype = None
        L.sgs_sample_fill_host.argtypes = [vp, u64, u64, vp, C.c_int]
        L.sgs_sample_fill_device.argtypes = [vp, C.c_int, u64, u64, vp, vp]
        L.sgs_sample_write_fastq.argtypes = [vp, u64, u64, C.c_char_p, u64, C.c_int, C.c_int]
        L.sgs_sample_write_fastq_bgzf.argtypes = [vp, u64, u64, C.c_char_p, C.c_int, C.c_int, u32]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(_load().sgs_last_error().decode())


def make_library(seed: int, n: int, k: int) -> np.ndarray:
    """uint8[n, k] ASCII guides"""
    out = np.empty((n, k), dtype=np.uint8)
    _check(_load().sgs_make_library(seed, n, k, out.ctypes.data))
    return out


class Sample:
    """One synthetic sample: fixed prefix/offset/orientation and a log-normal guide abundance."""

    def __init__(self, seed: int, sample_idx: int, library: np.ndarray, read_len: int, offset: int,
                 reverse: bool = False):
        library = np.ascontiguousarray(library, dtype=np.uint8)
        self.n, self.k = library.shape
        self.read_len, self.offset, self.reverse = read_len, offset, reverse
        self.stride = read_len + 1
        self._h = C.c_void_p()
        _check(_load().sgs_sample_create(seed, sample_idx, library.ctypes.data, self.n, self.k, read_len, offset,
                                         int(reverse), C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _load().sgs_sample_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def fill_host(self, first: int, n_reads: int, out: Optional[np.ndarray] = None, n_threads: int = 0) -> np.ndarray:
        if out is None:
            out = np.empty(n_reads * self.stride, dtype=np.uint8)
        assert out.nbytes >= n_reads * self.stride
        _check(_load().sgs_sample_fill_host(self._h, first, n_reads, out.ctypes.data, n_threads or (os.cpu_count() or 1)))
        return out

    def fill_host_ptr(self, first: int, n_reads: int, ptr: int, n_threads: int = 0) -> None:
        _check(_load().sgs_sample_fill_host(self._h, first, n_reads, ptr, n_threads or (os.cpu_count() or 1)))

    def fill_device(self, first: int, n_reads: int, d_ptr: int, device: int = 0, stream: Optional[int] = None) -> None:
        _check(_load().sgs_sample_fill_device(self._h, device, first, n_reads, d_ptr, stream))

    def write_fastq(self, path: str, first: int, n_reads: int, reads_per_member: int = 1 << 20, gz_level: int = 1,
                    n_threads: int = 0) -> None:
        _check(_load().sgs_sample_write_fastq(self._h, first, n_reads, path.encode(), reads_per_member, gz_level,
                                              n_threads or (os.cpu_count() or 1)))

    def write_fastq_bgzf(self, path: str, first: int, n_reads: int, gz_level: int = 1, n_threads: int = 0,
                         block_bytes: int = 65280) -> None:
        """BGZF (bgzip's blocked gzip): <= 64 KB members with the 'BC' extra field, cut anywhere"""
        _check(_load().sgs_sample_write_fastq_bgzf(self._h, first, n_reads, path.encode(), gz_level,
                                                   n_threads or (os.cpu_count() or 1), block_bytes))
