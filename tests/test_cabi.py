"""CPU: the C-ABI shared library loads and exports every symbol include/mergerec_b200.h declares (no compute
calls without a GPU), and the ctypes table in mergerec_b200/_lib.py covers exactly those symbols."""
import ctypes
import os
import re

import pytest

from mergerec_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mergerec_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    names = declared_functions()
    assert "mr_merge_axpy" in names and "mr_version" in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES.keys()) == declared_functions()


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.mr_version() >= 100
    assert isinstance(lib.mr_last_error(), bytes)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MergeRecLibraryError):
        _lib.require_cuda()


def test_argument_errors_do_not_need_a_gpu():
    """Argument validation happens before any CUDA call: NULL pointers -> MR_ERR_INVALID_ARG + message."""
    lib = _lib.load()
    rc = lib.mr_merge_axpy(None, None, 3, 16, None, 1, None, None, 1, 0, 1, None, None)
    assert rc == -1 and b"null" in lib.mr_last_error()
