"""Synthetic "bge-m3-shaped" chunk embeddings -- host (NumPy) side.

SURVEY.md section 8(d): dim 1024, unit norm, anisotropic and clustered:
``x_i = normalize(0.5*m + 0.6*c[cid(i)] + 0.62*g_i)`` with one common mean
direction ``m``, ``C`` cluster centres ``c_j`` and per-row noise ``g_i``; queries are
``normalize(x_r + 0.5*g')`` for a pseudo-random stored row ``r``.  The stand-in for
the remote bge-m3 embedding service of the reference
(reference app/llm_services.py:218-222) -- there is no network here.

The recipe is COUNTER-BASED and uses only exactly-rounded IEEE operations in a
fixed order (integer hash -> Irwin-Hall(4) of 16-bit fields -> fp32 mul/add ->
binary64 halving-tree norm -> fp32 scale), so that the CUDA generator in
``orx_testkit/csrc/synth.cu`` (``orxtk_synth_rows``) produces the SAME BITS for any row index.
That lets a 10M-row table be generated on the device while the host regenerates
any individual row for sampled verification, and lets the CPU reference arm of
``bench.py`` run on identical data.  tests/test_synth.py checks host == device.
"""
from __future__ import annotations

import numpy as np

DIM = 1024
SEED_TABLE = 20261018     # table stream
SEED_QUERY = 20261019     # query stream
SEED_CHURN = 20261020     # mixed-workload doc sizes

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_MASK16 = np.uint64(0xFFFF)

# stream tags (xor-ed into the seed before the first mix)
TAG_MEAN, TAG_CENTRE, TAG_NOISE, TAG_CID, TAG_QROW, TAG_QNOISE = (
    np.uint64(t) for t in (0x6D65616E, 0x63656E74, 0x6E6F6973, 0x63696421, 0x71726F77, 0x716E6F69))

GSCALE = np.float32(np.sqrt(3.0) / 65536.0)   # Irwin-Hall(4 x u16) -> unit variance
K_MEAN = np.float32(0.5)
K_CENTRE = np.float32(0.6)
K_NOISE = np.float32(0.62 / 32.0)             # noise vector has norm ~32 -> 0.62 * unit
K_QNOISE = np.float32(0.5 / 32.0)


def mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def stream_key(seed: int, tag: np.uint64) -> np.uint64:
    return mix64(np.uint64(seed) ^ tag)[()]


def _gauss(key: np.uint64, vec_index: np.ndarray) -> np.ndarray:
    """fp32 [len(vec_index), DIM] approximately N(0,1) elements of vectors ``vec_index``."""
    vec_index = np.asarray(vec_index, dtype=np.uint64).reshape(-1, 1)
    d = np.arange(DIM, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        ctr = vec_index * np.uint64(DIM) + d
        h = mix64(key + ctr * _GOLD)
    s = (h & _MASK16) + ((h >> np.uint64(16)) & _MASK16) + ((h >> np.uint64(32)) & _MASK16) + (h >> np.uint64(48))
    return (s.astype(np.int64) - 131070).astype(np.float32) * GSCALE


def _canon_sum(a: np.ndarray) -> np.ndarray:
    n = a.shape[-1]
    while n > 1:
        n //= 2
        a = a[..., :n] + a[..., n:2 * n]
    return a[..., 0]


def _normalize(y32: np.ndarray) -> np.ndarray:
    y = y32.astype(np.float64)
    n2 = _canon_sum(y * y)
    inv = (1.0 / np.sqrt(n2)).astype(np.float32)
    return y32 * inv[..., None]


class Synth:
    """Generator state: the common mean and the centre table (built once)."""

    def __init__(self, n_centres: int, seed: int = SEED_TABLE):
        self.seed = int(seed)
        self.n_centres = int(n_centres)
        self.k_mean = stream_key(seed, TAG_MEAN)
        self.k_centre = stream_key(seed, TAG_CENTRE)
        self.k_noise = stream_key(seed, TAG_NOISE)
        self.k_cid = stream_key(seed, TAG_CID)
        self.mean = _normalize(_gauss(self.k_mean, np.zeros(1, np.uint64)))[0]
        self.centres = np.empty((self.n_centres, DIM), np.float32)
        for s in range(0, self.n_centres, 4096):
            e = min(s + 4096, self.n_centres)
            self.centres[s:e] = _normalize(_gauss(self.k_centre, np.arange(s, e, dtype=np.uint64)))

    def cluster_of(self, rows: np.ndarray) -> np.ndarray:
        rows = np.asarray(rows, dtype=np.uint64)
        with np.errstate(over="ignore"):
            return (mix64(self.k_cid + rows * _GOLD) % np.uint64(self.n_centres)).astype(np.int64)

    def rows(self, index: np.ndarray) -> np.ndarray:
        """fp32 [len(index), DIM] embeddings of the rows with the given indices."""
        index = np.asarray(index, dtype=np.uint64)
        out = np.empty((index.shape[0], DIM), np.float32)
        for s in range(0, index.shape[0], 8192):
            idx = index[s:s + 8192]
            c = self.centres[self.cluster_of(idx)]
            g = _gauss(self.k_noise, idx)
            y = (K_MEAN * self.mean[None, :] + K_CENTRE * c) + K_NOISE * g
            out[s:s + 8192] = _normalize(y)
        return out

    def table(self, n_rows: int, start: int = 0) -> np.ndarray:
        return self.rows(np.arange(start, start + n_rows, dtype=np.uint64))

    def queries(self, n_queries: int, n_rows: int, seed: int = SEED_QUERY, start: int = 0):
        """``(q fp32 [n_queries, DIM], anchor_row int64 [n_queries])``; batches are prefixes."""
        k_row = stream_key(seed, TAG_QROW)
        k_noise = stream_key(seed, TAG_QNOISE)
        j = np.arange(start, start + n_queries, dtype=np.uint64)
        with np.errstate(over="ignore"):
            anchor = mix64(k_row + j * _GOLD) % np.uint64(n_rows)
        x = self.rows(anchor)
        y = x + K_QNOISE * _gauss(k_noise, j)
        return _normalize(y), anchor.astype(np.int64)


def default_centres(n_rows: int) -> int:
    """1024 centres up to 100k rows, 16384 from 1M rows on."""
    return 1024 if n_rows < 1_000_000 else 16384


def doc_chunk_counts(n_docs: int, seed: int = SEED_CHURN) -> np.ndarray:
    """Chunks per document for the mixed workload: 8..40, mean ~20
    (reference app/rag.py:112-116 splits at 1024 chars)."""
    j = np.arange(n_docs, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = mix64(stream_key(seed, np.uint64(0x646F6373)) + j * _GOLD)
    # triangular-ish on [8, 40] with mean ~20: min of two uniforms stretches the low end
    a = (h & np.uint64(0xFFFFFFFF)).astype(np.float64) / 2**32
    b = (h >> np.uint64(32)).astype(np.float64) / 2**32
    u = np.minimum(a, b) * 0.75 + np.maximum(a, b) * 0.25
    return (8 + np.floor(u * 33)).astype(np.int64).clip(8, 40)
