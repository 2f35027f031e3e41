"""CPU oracle for the Outline-RAG retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
(``outline_rag_b200``) never does; it fails loudly when the CUDA library is
missing.

PARITY UNPINNED.  The reference (/root/reference) contains no similarity
arithmetic, no tests, no fixtures and no golden vectors for this path
(SURVEY.md section 4 / 8c).  The arithmetic lives in two un-vendored third-party
dependencies that are absent from this image:

* ``langchain-postgres==0.0.16`` (reference requirements.txt:14) --
  ``AsyncPGVectorStore.__query_collection`` emits
  ``SELECT ..., cosine_distance("embedding", :q) AS distance FROM
  "langchain_pg_embedding" ORDER BY "embedding" <=> :q LIMIT :k``
  (reached from reference app/rag.py:69-79 (create), :85-87 (as_retriever, k=TOP_K),
  called at app/blueprints/api.py:122).
* pgvector (Docker tag ``pgvector/pgvector:pg16``, version not pinned,
  reference README.md:139; ``CREATE EXTENSION vector`` at app/database.py:66) --
  ``src/vector.c: cosine_distance``: three accumulators (dot, |a|^2, |b|^2) over
  the fp32 elements, ``similarity = dot / sqrt(norma * normb)`` finished in
  double, clamped to [-1, 1], ``distance = 1 - similarity``; a zero-norm operand
  gives NaN, which Postgres orders after every number.

This module restates that published algorithm.  What pins it instead of
reference fixtures are the eight known-answer constructions of SURVEY.md section 4
(tests/test_oracle.py, tests/golden/).

Two precisions are provided:

``canon_*``  (oracle-64)  every product and sum in IEEE binary64 in ONE FIXED
    ORDER (a halving tree, see :func:`canon_sum`).  This is the ground truth and
    DEFINES the canonical distance: the CUDA ``rescore`` kernel evaluates the
    identical tree, so ids AND distances can be compared bit for bit.
``pgv_*``    (oracle-pgv) fp32 accumulators and a double finish, i.e. pgvector's
    own precision.  Queries on which the two disagree are fp32-ambiguous
    (near-ties below fp32 resolution), not failures.

Ordering contract (SQL ``ORDER BY distance LIMIT k``; ties are unspecified by
SQL, this build defines them): ``(distance ASC, NaN last, id ASC)`` where ids
are 128-bit UUIDs compared as big-endian integers (= Postgres uuid order).
"""
from __future__ import annotations

import numpy as np

DIM = 1024

# --------------------------------------------------------------------------- ids


def ids_from_ints(values) -> np.ndarray:
    """128-bit ids as a uint64 [n, 2] array of (hi, lo) words.

    ``uuid.UUID(int=v)`` <-> (v >> 64, v & (2**64-1)); integer order equals the
    byte order Postgres uses for ``uuid`` (reference app/database.py:119,
    ``langchain_id UUID PRIMARY KEY``).
    """
    out = np.empty((len(values), 2), dtype=np.uint64)
    for i, v in enumerate(values):
        v = int(v)
        out[i, 0] = (v >> 64) & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = v & 0xFFFFFFFFFFFFFFFF
    return out


def ids_arange(start: int, stop: int) -> np.ndarray:
    """ids start..stop-1 (all < 2**64) without a Python loop."""
    out = np.zeros((stop - start, 2), dtype=np.uint64)
    out[:, 1] = np.arange(start, stop, dtype=np.uint64)
    return out


def ids_to_ints(ids: np.ndarray) -> list[int]:
    return [(int(h) << 64) | int(l) for h, l in np.asarray(ids, dtype=np.uint64).reshape(-1, 2)]


# ------------------------------------------------------------ canonical arithmetic


def canon_sum(a: np.ndarray) -> np.ndarray:
    """Fixed-order binary64 sum over the last axis (length a power of two).

    Halving tree: level by level ``a[i] <- a[i] + a[i + n/2]`` for ``i < n/2``.
    Each level is one IEEE add per element, so NumPy and the GPU (which keeps
    element ``i`` in lane ``i % 32`` and finishes with ``shfl_down`` 16,8,4,2,1)
    produce the same bits.
    """
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[-1]
    assert n & (n - 1) == 0 and n > 0, "canon_sum needs a power-of-two length"
    while n > 1:
        n //= 2
        a = a[..., :n] + a[..., n:2 * n]
    return a[..., 0]


def canon_sqnorm(X32: np.ndarray) -> np.ndarray:
    X = np.asarray(X32, dtype=np.float32).astype(np.float64)
    return canon_sum(X * X)  # fp32*fp32 products are exact in binary64


def canon_distance(X32: np.ndarray, q32: np.ndarray, chunk: int = 8192) -> np.ndarray:
    """Canonical cosine distance of every row of ``X32`` [n, d] to ``q32`` [d].

    Follows pgvector ``cosine_distance`` [UPSTREAM] with binary64 accumulators:
    ``1 - clamp(dot / sqrt(|x|^2 * |q|^2), -1, 1)``; NaN when a norm is zero.
    """
    X32 = np.asarray(X32, dtype=np.float32)
    q = np.asarray(q32, dtype=np.float32).astype(np.float64)
    n2q = canon_sum(q * q)
    out = np.empty(X32.shape[0], dtype=np.float64)
    for s in range(0, X32.shape[0], chunk):
        X = X32[s:s + chunk].astype(np.float64)
        dot = canon_sum(X * q)
        n2x = canon_sum(X * X)
        with np.errstate(invalid="ignore", divide="ignore"):
            sim = dot / np.sqrt(n2x * n2q)
        # pgvector clamps to [-1, 1]; NaN stays NaN through the comparisons
        sim = np.where(sim > 1.0, 1.0, sim)
        sim = np.where(sim < -1.0, -1.0, sim)
        out[s:s + chunk] = 1.0 - sim
    return out


def order_by_distance(dist: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Permutation implementing ``ORDER BY distance ASC`` (NaN last), id ASC on ties."""
    ids = np.asarray(ids, dtype=np.uint64).reshape(-1, 2)
    nan = np.isnan(dist)
    d = np.where(nan, 0.0, dist)
    # lexsort: last key is the primary one
    return np.lexsort((ids[:, 1], ids[:, 0], d, nan))


def validate_vectors(V: np.ndarray, dim: int = DIM) -> None:
    """pgvector input checks [UPSTREAM vector.c CheckDim/CheckElement]:
    wrong dimension and NaN/Inf elements are errors, not data."""
    V = np.asarray(V)
    if V.ndim != 2 or V.shape[1] != dim:
        raise ValueError(f"expected {dim} dimensions, not {V.shape[-1] if V.ndim else 0}")
    if not np.isfinite(V).all():
        raise ValueError("NaN or infinite value not allowed in vector")


def topk_exact(X32, ids, q32, k: int, exhaustive: bool = False):
    """Ground-truth top-k: ``(ids [m,2] uint64, distance [m] float64)``, m = min(k, n).

    ``exhaustive=True`` evaluates the canonical tree on every row (small n).
    Otherwise a binary64 BLAS pass shortlists every row whose approximate
    distance is within 1e-9 of the k-th (BLAS vs canonical differ by < 1e-12 for
    unit-scale data), and only the shortlist is evaluated canonically -- still
    exact, the margin is three orders above the error bound.
    """
    X32 = np.asarray(X32, dtype=np.float32)
    ids = np.asarray(ids, dtype=np.uint64).reshape(-1, 2)
    q32 = np.asarray(q32, dtype=np.float32)
    n = X32.shape[0]
    if n == 0 or k <= 0:
        return np.zeros((0, 2), np.uint64), np.zeros(0, np.float64)
    if exhaustive or n <= 4 * k + 64:
        rows = np.arange(n)
    else:
        q = q32.astype(np.float64)
        n2q = float(q @ q)
        approx = np.empty(n, np.float64)
        for s in range(0, n, 65536):
            X = X32[s:s + 65536].astype(np.float64)
            with np.errstate(invalid="ignore", divide="ignore"):
                approx[s:s + 65536] = 1.0 - (X @ q) / np.sqrt(np.einsum("ij,ij->i", X, X) * n2q)
        nan = np.isnan(approx)
        finite = np.where(nan, np.inf, approx)
        kth = np.partition(finite, min(k, n) - 1)[min(k, n) - 1]
        if np.isinf(kth):            # fewer than k finite rows: everything is a candidate
            rows = np.arange(n)
        else:
            rows = np.nonzero(finite <= kth + 1e-9)[0]
    d = canon_distance(X32[rows], q32)
    perm = order_by_distance(d, ids[rows])[:k]
    return ids[rows][perm].copy(), d[perm].copy()


class StreamingTopK:
    """Ground-truth top-k of MANY rows fed chunk by chunk (the full-scan check of bench.py / the 1M-row
    tests: a 10M x 1024 table does not fit one NumPy pass, and `topk_exact` converts to binary64).

    Per chunk one fp32 BLAS pass (`X @ Q.T`, all queries at once) shortlists, for every query, each row whose
    approximate cosine is within `margin` of the running k-th best; the shortlist keeps the rows' vectors, and
    `result()` evaluates the CANONICAL distance (`canon_distance`) on it and applies the ordering contract.
    Exactness: |fp32 BLAS cosine - exact cosine| <= ~dim * 2^-24 * sum|x_i q_i| / (|x||q|) <= 6.1e-5 for ANY
    summation order; `margin` = 2e-4 covers the error of both the candidate and the k-th, so no row of the
    true top-k is ever dropped.  Rows whose fp32 norm is zero / non-finite (distance NaN or untrusted) are
    always kept.  ``rows_are_bf16``: chunks are raw bf16 bit patterns (uint16) of a bf16 table.
    """

    def __init__(self, Q32, k: int, margin: float = 2e-4):
        self.Q = np.ascontiguousarray(np.asarray(Q32, dtype=np.float32).reshape(-1, DIM))
        self.k = int(k)
        self.margin = float(margin)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.inv_q = (1.0 / np.sqrt(np.einsum("ij,ij->i", self.Q, self.Q).astype(np.float64))).astype(np.float32)
        nq = self.Q.shape[0]
        self._vec = [np.zeros((0, DIM), np.float32) for _ in range(nq)]
        self._ids = [np.zeros((0, 2), np.uint64) for _ in range(nq)]
        self._sim = [np.zeros(0, np.float32) for _ in range(nq)]
        self.rows_seen = 0

    @staticmethod
    def bf16_bits_to_f32(raw_u16: np.ndarray) -> np.ndarray:
        return (np.ascontiguousarray(raw_u16, np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)

    def feed(self, X, ids, rows_are_bf16: bool = False) -> None:
        X = self.bf16_bits_to_f32(X) if rows_are_bf16 else np.ascontiguousarray(X, np.float32)
        ids = np.asarray(ids, dtype=np.uint64).reshape(-1, 2)
        n = X.shape[0]
        if n == 0:
            return
        self.rows_seen += n
        n2 = np.einsum("ij,ij->i", X, X)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv_x = (1.0 / np.sqrt(n2)).astype(np.float32)
        special = ~np.isfinite(inv_x) | ~(n2 > np.float32(1e-30)) | ~(n2 < np.float32(1e30))
        S = self.Q @ X.T                                      # [nq, n] fp32 BLAS (this orientation is ~10x faster
        with np.errstate(invalid="ignore"):                   #  in OpenBLAS than the skinny X @ Q.T)
            S *= inv_x[None, :]
            S *= self.inv_q[:, None]
        for j in range(self.Q.shape[0]):
            s = S[j]
            fin = np.where(special | ~np.isfinite(s), -np.inf, s)
            pool = np.concatenate([self._sim[j], fin])
            pool_fin = pool[np.isfinite(pool)]
            if pool_fin.shape[0] >= self.k:
                kth = np.partition(pool_fin, pool_fin.shape[0] - self.k)[pool_fin.shape[0] - self.k]
                thr = kth - np.float32(self.margin)
            else:
                thr = -np.inf
            keep_new = np.nonzero((fin >= thr) | special)[0]
            keep_old = np.nonzero((self._sim[j] >= thr) | ~np.isfinite(self._sim[j]))[0]
            self._vec[j] = np.concatenate([self._vec[j][keep_old], X[keep_new]])
            self._ids[j] = np.concatenate([self._ids[j][keep_old], ids[keep_new]])
            self._sim[j] = np.concatenate([self._sim[j][keep_old], np.where(special[keep_new], -np.inf, fin[keep_new])])

    def candidates(self, j: int):
        """(ids [m,2], canonical distance [m]) of query j's shortlist, unordered -- for cross-shard merges."""
        return self._ids[j].copy(), canon_distance(self._vec[j], self.Q[j])

    def result(self):
        """-> list over queries of (ids [m,2] uint64, distance [m] float64), m = min(k, rows)."""
        out = []
        for j in range(self.Q.shape[0]):
            ids, d = self.candidates(j)
            perm = order_by_distance(d, ids)[:self.k]
            out.append((ids[perm].copy(), d[perm].copy()))
        return out


# ------------------------------------------------------------ pgvector-precision path


def pgv_distance(X32: np.ndarray, q32: np.ndarray) -> np.ndarray:
    """oracle-pgv: fp32 accumulators (BLAS sgemv order), double finish [UPSTREAM]."""
    X32 = np.asarray(X32, dtype=np.float32)
    q32 = np.asarray(q32, dtype=np.float32)
    dot = X32 @ q32
    n2x = np.einsum("ij,ij->i", X32, X32)
    n2q = np.float32(q32 @ q32)
    with np.errstate(invalid="ignore", divide="ignore"):
        sim = dot.astype(np.float64) / np.sqrt(n2x.astype(np.float64) * np.float64(n2q))
    sim = np.where(sim > 1.0, 1.0, sim)
    sim = np.where(sim < -1.0, -1.0, sim)
    return 1.0 - sim


def numpy_replica_topk(X32, ids, q32, k: int, row_inv_norm=None):
    """The CPU baseline north_star mandates when Postgres cannot be installed:
    OpenBLAS ``sgemv`` + ``argpartition`` + ``lexsort((id, distance))``.

    ``row_inv_norm`` (fp32 [n]) may be passed when the table norms were
    precomputed once (Postgres recomputes them per query; passing them favours
    the CPU baseline and is what bench.py does).
    """
    X32 = np.asarray(X32, dtype=np.float32)
    ids = np.asarray(ids, dtype=np.uint64).reshape(-1, 2)
    q32 = np.asarray(q32, dtype=np.float32)
    n = X32.shape[0]
    if row_inv_norm is None:
        dist = pgv_distance(X32, q32)
    else:
        inv_q = 1.0 / np.sqrt(np.float64(q32 @ q32))
        sim = (X32 @ q32).astype(np.float64) * row_inv_norm * inv_q
        dist = 1.0 - np.clip(sim, -1.0, 1.0)
    m = min(k, n)
    if m == 0:
        return np.zeros((0, 2), np.uint64), np.zeros(0, np.float64)
    finite = np.where(np.isnan(dist), np.inf, dist)
    part = np.argpartition(finite, m - 1)[:m] if m < n else np.arange(n)
    # rows tied with the m-th distance outside the partition must compete on id
    kth = finite[part].max()
    if np.isfinite(kth):
        part = np.nonzero(finite <= kth)[0]
    else:
        part = np.arange(n)
    perm = order_by_distance(dist[part], ids[part])[:m]
    return ids[part][perm].copy(), dist[part][perm].copy()


# ------------------------------------------------------------------ shard merge


def merge_shards(parts, k: int):
    """Merge per-shard ``(ids, distance)`` top-k lists into the global top-k with
    the same ordering contract (what the allgather + device merge must equal)."""
    ids = np.concatenate([np.asarray(p[0], np.uint64).reshape(-1, 2) for p in parts], axis=0)
    dist = np.concatenate([np.asarray(p[1], np.float64).reshape(-1) for p in parts], axis=0)
    perm = order_by_distance(dist, ids)[:k]
    return ids[perm], dist[perm]


def recall_at_k(got_ids, want_ids) -> float:
    want = {tuple(map(int, r)) for r in np.asarray(want_ids).reshape(-1, 2)}
    if not want:
        return 1.0
    got = {tuple(map(int, r)) for r in np.asarray(got_ids).reshape(-1, 2)}
    return len(got & want) / len(want)
