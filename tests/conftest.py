"""Shared fixtures.  `-m "not gpu"` runs on the CPU container (oracle, host logic, ABI);
`-m gpu` tests are the parity tests proper and call through the C-ABI on a B200."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pgv_lib():
    """The C restatement of pgvector's scan loop (oracle/pgv_cosine.c), built on demand."""
    path = os.path.join(ROOT, "oracle", "libpgv_cosine.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(path)
    lib.pgv_cosine_distance.restype = ctypes.c_double
    lib.pgv_cosine_distance.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.pgv_scan_topk.restype = ctypes.c_int
    lib.pgv_scan_topk.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.pgv_scan_topk_mt.restype = ctypes.c_int
    lib.pgv_scan_topk_mt.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


@pytest.fixture(scope="session")
def synth100k():
    """Generator state for the 100k-row config (1024 centres), SURVEY.md 8(d)."""
    from orx_testkit.synth import Synth
    return Synth(1024)


@pytest.fixture(scope="session")
def small_table(synth100k):
    """8192 rows of the synthetic table + 32 queries (host, fp32)."""
    X = synth100k.table(8192)
    Q, anchors = synth100k.queries(32, 8192)
    return X, Q, anchors
