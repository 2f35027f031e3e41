"""ctypes binding of ``orx_testkit/liborx_synth.so`` (csrc/synth.cu): synthetic rows generated in HBM,
bit-identical to :class:`orx_testkit.synth.Synth`.  Test / bench tooling only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborx_synth.so")
_lib = None


def build() -> str:
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building liborx_synth.so failed:\n" + out.stdout + out.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make -C {_HERE}` (or __graft_entry__.build())")
        l = C.CDLL(LIB_PATH)
        l.orxtk_synth_rows.restype = C.c_int
        l.orxtk_synth_rows.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        l.orxtk_last_error.restype = C.c_char_p
        _lib = l
    return _lib


def synth_rows_device(device: int, seed: int, n_centres: int, row_start: int, n_rows: int, out=None):
    """Rows ``row_start .. row_start+n_rows-1`` of the synthetic table, generated in HBM on torch's
    current stream (bit-identical to ``synth.Synth.rows``).  Returns a float32 CUDA tensor."""
    import torch
    if out is None:
        out = torch.empty((n_rows, 1024), dtype=torch.float32, device=f"cuda:{device}")
    stream = torch.cuda.current_stream(device).cuda_stream
    rc = lib().orxtk_synth_rows(int(device), C.c_void_p(stream), int(seed), int(n_centres), int(row_start),
                                int(n_rows), C.c_void_p(out.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"orxtk_synth_rows failed ({rc}): {lib().orxtk_last_error().decode()}")
    return out
