"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/pbn_cuda.h
declares; the product path fails loudly (no CPU fallback) when no GPU is usable."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pbn_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from pybnesian_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the CUDA extension first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in include/pbn_cuda.h is not exported" % s
    # the Python binding lists exactly the header's entry points
    assert sorted(_lib.EXPORTS) == syms


def test_library_reports_version_and_errors_without_compute():
    from pybnesian_b200 import _lib
    L = _lib.lib()
    assert b"sm_100a" in L.pbn_version()
    n = ctypes.c_int(-1)
    rc = L.pbn_device_count(ctypes.byref(n))
    assert rc in (_lib.PBN_OK, _lib.PBN_ERR_CUDA)
    if rc != _lib.PBN_OK or n.value == 0:
        # no usable GPU here: the product must refuse, never fall back to a CPU path
        with pytest.raises((RuntimeError, ValueError)):
            _lib.Context(0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "pybnesian_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "pbn_oracle" not in text, f
