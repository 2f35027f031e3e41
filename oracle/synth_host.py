"""Fast host generator of the synthetic table (ctypes over oracle/libsynth_host.so, plain C + pthreads) --
TEST / BASELINE INFRASTRUCTURE ONLY, like everything under oracle/.

`FastSynth` is `orx_testkit.synth.Synth` with `rows()` / the centre table produced by the C generator, which is
bit-identical to the NumPy one (tests/test_synth.py) and ~100x faster, so the CPU arm of bench.py can time the
reference ordering on a MEASURED 1M-row table (BASELINE.json configs[1]) instead of a scaled-up 200k sample.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from orx_testkit.synth import DIM, SEED_TABLE, Synth

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libsynth_host.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE, "libsynth_host.so"], check=True, capture_output=True)
        lib = C.CDLL(_PATH)
        lib.synth_host_unit.restype = None
        lib.synth_host_unit.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        lib.synth_host_rows.restype = C.c_int
        lib.synth_host_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                        C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


class FastSynth(Synth):
    def __init__(self, n_centres: int, seed: int = SEED_TABLE, threads: int | None = None):
        self.threads = int(threads or len(os.sched_getaffinity(0)))
        lib = _load()
        # keys and the mean come from the NumPy class (one vector); the centre table from C
        self._centres_c = np.empty((int(n_centres), DIM), np.float32)
        Synth.__init__(self, 1, seed)                       # builds keys + mean (+ one throw-away centre)
        self.n_centres = int(n_centres)
        lib.synth_host_unit(int(self.k_centre), 0, self.n_centres, C.c_void_p(self._centres_c.ctypes.data))
        self.centres = self._centres_c

    def rows(self, index: np.ndarray) -> np.ndarray:
        index = np.ascontiguousarray(index, dtype=np.uint64)
        out = np.empty((index.shape[0], DIM), np.float32)
        if index.shape[0]:
            rc = _load().synth_host_rows(int(self.k_noise), int(self.k_cid), C.c_void_p(self.mean.ctypes.data),
                                         C.c_void_p(self.centres.ctypes.data), self.n_centres,
                                         C.c_void_p(index.ctypes.data), 0, index.shape[0],
                                         C.c_void_p(out.ctypes.data), self.threads)
            if rc != 0:
                raise RuntimeError(f"synth_host_rows failed ({rc})")
        return out

    def table(self, n_rows: int, start: int = 0) -> np.ndarray:
        out = np.empty((int(n_rows), DIM), np.float32)
        if n_rows:
            rc = _load().synth_host_rows(int(self.k_noise), int(self.k_cid), C.c_void_p(self.mean.ctypes.data),
                                         C.c_void_p(self.centres.ctypes.data), self.n_centres, None, int(start),
                                         int(n_rows), C.c_void_p(out.ctypes.data), self.threads)
            if rc != 0:
                raise RuntimeError(f"synth_host_rows failed ({rc})")
        return out
