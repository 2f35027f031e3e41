"""The C-ABI library loads without a GPU and exports every symbol include/orx.h declares.
No compute calls here; `orx_create` must fail LOUDLY (no CPU fallback) when no device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "orx.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orx_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = _declared_symbols()
    for must in ["orx_create", "orx_destroy", "orx_upsert", "orx_delete", "orx_search", "orx_merge_topk",
                 "orx_size", "orx_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from outline_rag_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/orx.h but not exported by liborx.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(_declared_symbols())


def test_struct_layouts_match_header():
    from outline_rag_b200 import _lib
    assert ctypes.sizeof(_lib.OrxId) == 16
    assert ctypes.sizeof(_lib.OrxStats) == 6 * 8 + 2 * 4 + 2 * 4 + 8 + 8


def test_version_and_argument_errors_need_no_gpu():
    from outline_rag_b200 import _lib
    assert "sm_100a" in _lib.version()
    h = ctypes.c_void_p()
    assert _lib.lib.orx_create(ctypes.byref(h), 768, 0, 0, 0) == _lib.ORX_ERR_DIM
    assert "expected 1024 dimensions" in _lib.last_error()
    assert _lib.lib.orx_create(ctypes.byref(h), 1024, 7, 0, 0) == _lib.ORX_ERR_INVALID
    assert _lib.lib.orx_search(None, None, 1, 1024, 12, None, None, None) == _lib.ORX_ERR_INVALID


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import outline_rag_b200 as orx
    with pytest.raises(orx.OrxError, match="no CPU fallback"):
        orx.Index("fp32")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "outline_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
