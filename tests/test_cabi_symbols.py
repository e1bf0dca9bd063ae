"""CPU-side checks of the drop-in boundary: the shared object loads and exports exactly the
symbols include/sgcount_cuda.h declares, and it fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

from sgcount_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sgcount_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(sgc_[a-z_0-9]+)\s*\(", text))


def test_header_and_binding_agree():
    assert header_symbols() == set(_cabi.SIGNATURES)


def test_library_exports_every_symbol():
    lib = _cabi.load()
    for name in header_symbols():
        assert getattr(lib, name) is not None
    assert lib.sgc_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu():
    lib = _cabi.load()
    n = C.c_int(-1)
    rc = lib.sgc_device_count(C.byref(n))
    if rc == _cabi.OK and n.value > 0:
        pytest.skip("a CUDA device is present")
    # with no device every entry point must fail with SGC_ERR_CUDA, never compute on the CPU
    out = C.c_void_p()
    seq = C.create_string_buffer(b"ACTG")
    rc = lib.sgc_library_create(0, C.cast(seq, C.c_void_p), 1, 4, 1, C.byref(out))
    assert rc == _cabi.ERR_CUDA
    assert lib.sgc_last_error()
    assert not out.value


def test_product_does_not_import_the_oracle():
    """nothing under sgcount_b200/ may reference oracle/ (the oracle is the checker only)"""
    pkg = os.path.join(ROOT, "sgcount_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f
