"""Pins the oracle (oracle/cosine_topk.py, oracle/pgv_cosine.c).

The reference holds no tests, fixtures or golden vectors for this path (SURVEY.md 4 / 8c:
"parity unpinned"), so the pins are the eight known-answer constructions derived from the SQL
semantics of `ORDER BY embedding <=> :q LIMIT :k` (reference app/rag.py:85-87) and pgvector's
published `cosine_distance`, plus the committed golden vectors in tests/golden/.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import cosine_topk as O

DIM = 1024
K = 12


def _ids(n, start=0):
    return O.ids_arange(start, start + n)


# ---------------------------------------------------------------- known answers
def test_ka1_identity_basis():
    X = np.eye(DIM, dtype=np.float32)
    q = np.arange(DIM, 0, -1).astype(np.float32)
    ids, dist = O.topk_exact(X, _ids(DIM), q, K, exhaustive=True)
    assert O.ids_to_ints(ids) == list(range(K))
    want = 1.0 - (DIM - np.arange(K)) / np.sqrt(np.sum(q.astype(np.float64) ** 2))
    np.testing.assert_allclose(dist, want, rtol=0, atol=1e-15)


def test_ka2_duplicates_order_by_id():
    rng = np.random.default_rng(1)
    base = rng.standard_normal((4, DIM)).astype(np.float32)
    X = np.concatenate([base, base, base])          # every row three times
    ids = O.ids_from_ints([50, 40, 30, 20, 11, 12, 13, 14, 5, 6, 7, 8])
    got, dist = O.topk_exact(X, ids, base[0], 3, exhaustive=True)
    assert O.ids_to_ints(got) == [5, 11, 50]         # same row: ids ascending
    assert dist[0] == dist[1] == dist[2]


def test_ka3_scale_invariance():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((300, DIM)).astype(np.float32)
    q = rng.standard_normal(DIM).astype(np.float32)
    s = (2.0 ** rng.integers(-6, 7, size=300)).astype(np.float32)   # exact power-of-two scalings
    a_ids, a_d = O.topk_exact(X, _ids(300), q, K, exhaustive=True)
    b_ids, b_d = O.topk_exact(X * s[:, None], _ids(300), q, K, exhaustive=True)
    assert np.array_equal(a_ids, b_ids)
    np.testing.assert_array_equal(a_d, b_d)          # power-of-two scaling is exact in binary64 too


def test_ka4_antiparallel_is_two_and_clamped():
    rng = np.random.default_rng(3)
    q = rng.standard_normal(DIM).astype(np.float32)
    X = np.stack([q, -q, 4 * q])     # power-of-two scaling is exact
    d = O.canon_distance(X, q)
    assert d[0] == 0.0 and d[2] == 0.0 and d[1] == 2.0
    assert (d >= 0.0).all() and (d <= 2.0).all()


def test_ka5_zero_norm_row_sorts_last():
    rng = np.random.default_rng(4)
    X = rng.standard_normal((5, DIM)).astype(np.float32)
    X[1] = 0.0
    q = rng.standard_normal(DIM).astype(np.float32)
    ids, dist = O.topk_exact(X, _ids(5), q, 12, exhaustive=True)
    assert len(dist) == 5 and np.isnan(dist[-1]) and not np.isnan(dist[:-1]).any()
    assert O.ids_to_ints(ids)[-1] == 1
    ids4, dist4 = O.topk_exact(X, _ids(5), q, 4, exhaustive=True)
    assert 1 not in O.ids_to_ints(ids4)              # returned only when N_live < k


def test_ka6_k_larger_than_table():
    rng = np.random.default_rng(5)
    X = rng.standard_normal((7, DIM)).astype(np.float32)
    ids, dist = O.topk_exact(X, _ids(7), X[3], 12)
    assert len(dist) == 7 and O.ids_to_ints(ids)[0] == 3
    assert (np.diff(dist) >= 0).all()
    e_ids, e_d = O.topk_exact(X[:0], _ids(0), X[3], 12)
    assert e_ids.shape == (0, 2) and e_d.shape == (0,)


def test_ka8_input_validation():
    with pytest.raises(ValueError, match="expected 1024 dimensions"):
        O.validate_vectors(np.zeros((2, 768), np.float32))
    bad = np.zeros((2, DIM), np.float32)
    bad[1, 7] = np.nan
    with pytest.raises(ValueError, match="NaN or infinite"):
        O.validate_vectors(bad)
    bad[1, 7] = np.inf
    with pytest.raises(ValueError):
        O.validate_vectors(bad)


# ---------------------------------------------------------------- internal consistency
def test_canon_sum_matches_exact_rational():
    from fractions import Fraction
    rng = np.random.default_rng(6)
    a = rng.standard_normal(1024)
    exact = float(sum(Fraction(float(v)) for v in a))
    assert abs(O.canon_sum(a) - exact) <= 1e-12
    # fixed order: same bits on a second evaluation and for a 2-D batch
    assert O.canon_sum(a) == O.canon_sum(np.stack([a, a]))[1]


def test_shortlist_path_equals_exhaustive(small_table):
    X, Q, _ = small_table
    ids = _ids(X.shape[0])
    for q in Q[:8]:
        a = O.topk_exact(X, ids, q, K)
        b = O.topk_exact(X, ids, q, K, exhaustive=True)
        assert np.array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])


def test_pgv_precision_agrees_with_canonical(small_table):
    X, Q, _ = small_table
    ids = _ids(X.shape[0])
    for q in Q[:8]:
        c_ids, c_d = O.topk_exact(X, ids, q, K)
        p_ids, p_d = O.numpy_replica_topk(X, ids, q, K)
        np.testing.assert_allclose(p_d, c_d, rtol=0, atol=5e-7)    # fp32 accumulators
        # fp32-ambiguous near-ties may swap neighbours; the SET is stable on this data
        assert set(O.ids_to_ints(p_ids)) == set(O.ids_to_ints(c_ids))


def test_c_restatement_matches_numpy(pgv_lib, small_table):
    X, Q, _ = small_table
    n = X.shape[0]
    ids = _ids(n)
    rows = np.zeros(K, np.int64)
    dist = np.zeros(K, np.float64)
    for q in Q[:4]:
        q = np.ascontiguousarray(q)
        m = pgv_lib.pgv_scan_topk(X.ctypes.data, 0, n, DIM, q.ctypes.data, K, rows.ctypes.data, dist.ctypes.data)
        assert m == K
        c_ids, c_d = O.topk_exact(X, ids, q, K)
        np.testing.assert_allclose(dist, c_d, rtol=0, atol=5e-7)
        assert set(rows.tolist()) == set(O.ids_to_ints(c_ids))
        rows_mt = np.zeros(K, np.int64)
        dist_mt = np.zeros(K, np.float64)
        m = pgv_lib.pgv_scan_topk_mt(X.ctypes.data, n, DIM, q.ctypes.data, K, 4, rows_mt.ctypes.data,
                                     dist_mt.ctypes.data)
        assert m == K and np.array_equal(rows_mt, rows) and np.array_equal(dist_mt, dist)
    d1 = pgv_lib.pgv_cosine_distance(DIM, X[0].ctypes.data, X[0].ctypes.data)
    assert abs(d1) < 1e-6


def test_merge_shards_equals_global(small_table):
    X, Q, _ = small_table
    ids = _ids(X.shape[0])
    q = Q[0]
    parts = []
    for s in range(4):
        sel = np.arange(s, X.shape[0], 4)
        parts.append(O.topk_exact(X[sel], ids[sel], q, K))
    g = O.topk_exact(X, ids, q, K)
    m = O.merge_shards(parts, K)
    assert np.array_equal(m[0], g[0])
    np.testing.assert_array_equal(m[1], g[1])


# ---------------------------------------------------------------- committed golden vectors
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "topk_golden.json")


def test_golden_vectors(synth100k):
    """tests/golden/topk_golden.json (made by tests/golden/make_golden.py with THIS oracle on
    the counter-based synthetic table): guards the oracle and the generator against drift."""
    with open(GOLDEN) as f:
        G = json.load(f)
    X = synth100k.table(G["n_rows"])
    ids = _ids(G["n_rows"])
    Q, _ = synth100k.queries(len(G["cases"]), G["n_rows"])
    assert float(X[123, 45]).hex() == G["probe_x_123_45"]
    for qi, case in enumerate(G["cases"]):
        got_ids, got_d = O.topk_exact(X, ids, Q[qi], G["k"])
        assert O.ids_to_ints(got_ids) == case["ids"]
        assert [float(d).hex() for d in got_d] == case["dist_hex"]


# ---------------------------------------------------------------- properties (hypothesis, CPU)
from hypothesis import given, settings            # noqa: E402
from hypothesis import strategies as st           # noqa: E402


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 60), st.integers(1, 5), st.integers(1, 20))
def test_merging_any_partition_equals_the_global_answer(seed, n, parts, k):
    """ORDER BY ... LIMIT k over a union of disjoint shards == merge of the per-shard answers, for any
    partition (the invariant the row-sharded path relies on), including duplicates and zero rows."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((max(1, n // 3), DIM)).astype(np.float32)
    X = base[rng.integers(0, base.shape[0], size=n)]                 # many exact duplicates
    X[rng.random(n) < 0.1] = 0.0                                     # some zero-norm rows
    ids = O.ids_from_ints(rng.choice(10 * n + 10, size=n, replace=False).tolist())
    q = rng.standard_normal(DIM).astype(np.float32)
    owner = rng.integers(0, parts, size=n)
    shard_answers = [O.topk_exact(X[owner == p], ids[owner == p], q, k, exhaustive=True) for p in range(parts)]
    m_ids, m_d = O.merge_shards(shard_answers, k)
    g_ids, g_d = O.topk_exact(X, ids, q, k, exhaustive=True)
    assert np.array_equal(m_ids, g_ids)
    assert np.array_equal(np.isnan(m_d), np.isnan(g_d)) and np.array_equal(m_d[~np.isnan(m_d)], g_d[~np.isnan(g_d)])


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2**31 - 1))
def test_canonical_distance_is_within_an_ulp_of_exact_arithmetic(seed):
    import math
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(DIM).astype(np.float32)
    q = rng.standard_normal(DIM).astype(np.float32)
    d = O.canon_distance(x[None, :], q)[0]
    xd, qd = x.astype(np.float64), q.astype(np.float64)
    exact = 1.0 - math.fsum(xd * qd) / math.sqrt(math.fsum(xd * xd) * math.fsum(qd * qd))
    assert abs(d - exact) <= 1e-14
    assert O.canon_distance(x[None, :] * np.float32(4.0), q * np.float32(0.5))[0] == d      # scale-free, exactly


# ---------------------------------------------------------------- streaming full scan (bench.py verify, 1M-row tests)
def test_streaming_full_scan_equals_topk_exact(small_table):
    """`StreamingTopK` (fp32 BLAS shortlist per chunk, canonical rescoring of the shortlist) is what checks the
    engine at sizes `topk_exact` cannot take; it must give the same ids and distance bits, also with zero-norm,
    tiny-norm and duplicated rows in the table and when rows arrive as raw bf16."""
    from tests._helpers import bf16_rne
    X, Q, _ = small_table
    X = X.copy()
    X[100] = 0.0
    X[200] *= np.float32(1e-25)
    X[300] = X[7]
    ids = _ids(X.shape[0], start=2**40)
    st = O.StreamingTopK(Q[:6], K)
    for s in range(0, X.shape[0], 3000):
        st.feed(X[s:s + 3000], ids[s:s + 3000])
    assert st.rows_seen == X.shape[0]
    for j, (g_ids, g_d) in enumerate(st.result()):
        w_ids, w_d = O.topk_exact(X, ids, Q[j], K)
        assert np.array_equal(g_ids, w_ids) and np.array_equal(g_d.view(np.uint64), w_d.view(np.uint64)), j
    # bf16 rows fed as raw bit patterns
    Xb = bf16_rne(X)
    raw = (Xb.view(np.uint32) >> np.uint32(16)).astype(np.uint16)
    st = O.StreamingTopK(Q[:3], K)
    st.feed(raw, ids, rows_are_bf16=True)
    for j, (g_ids, g_d) in enumerate(st.result()):
        w_ids, w_d = O.topk_exact(Xb, ids, Q[j], K)
        assert np.array_equal(g_ids, w_ids) and np.array_equal(g_d.view(np.uint64), w_d.view(np.uint64)), j
    # fewer finite rows than k: NaN rows complete the answer in id order
    Xs = X[:8].copy()
    Xs[3] = 0.0
    st = O.StreamingTopK(Q[:1], K)
    st.feed(Xs, ids[:8])
    (g_ids, g_d), = st.result()
    w_ids, w_d = O.topk_exact(Xs, ids[:8], Q[0], K)
    assert np.array_equal(g_ids, w_ids) and np.array_equal(g_d.view(np.uint64), w_d.view(np.uint64))


def test_streaming_candidates_merge_across_shards(small_table):
    """What bench.py does at N > 1: per-shard `candidates()` merged by `merge_shards` == the global answer."""
    X, Q, _ = small_table
    ids = _ids(X.shape[0])
    parts = np.arange(X.shape[0]) % 3
    shards = []
    for p in range(3):
        st = O.StreamingTopK(Q[:4], K)
        st.feed(X[parts == p], ids[parts == p])
        shards.append([st.candidates(j) for j in range(4)])
    for j in range(4):
        m_ids, m_d = O.merge_shards([s[j] for s in shards], K)
        w_ids, w_d = O.topk_exact(X, ids, Q[j], K)
        assert np.array_equal(m_ids, w_ids) and np.array_equal(m_d.view(np.uint64), w_d.view(np.uint64))


def test_oracle_agrees_with_independent_third_party_implementations(small_table):
    """No reference fixture exists, but two independent implementations of the same published definition do: SciPy's
    `spatial.distance.cosine` and scikit-learn's brute-force cosine k-NN.  Distances agree to double rounding
    (tolerance 1e-12 stated here: they sum in another order than the canonical tree) and the neighbour ids are identical
    wherever the gap to the next distance exceeds that tolerance."""
    from scipy.spatial.distance import cosine as scipy_cosine
    from sklearn.neighbors import NearestNeighbors
    X, Q, _ = small_table
    X = X[:3000].astype(np.float64)
    ids = _ids(3000)
    nn = NearestNeighbors(n_neighbors=K + 1, metric="cosine", algorithm="brute").fit(X)
    nd, ni = nn.kneighbors(Q[:16].astype(np.float64))
    for i in range(16):
        w_ids, w_d = O.topk_exact(X.astype(np.float32), ids, Q[i], K)
        rows = [int(v) for v in O.ids_to_ints(w_ids)]
        for r, d in zip(rows, w_d):
            assert abs(scipy_cosine(X[r], Q[i].astype(np.float64)) - d) < 1e-12
        assert np.allclose(nd[i, :K], w_d, rtol=0, atol=1e-12)
        clear = np.diff(nd[i]) > 1e-10                           # positions whose successor is clearly farther
        for p in range(K):
            if clear[p] and (p == 0 or clear[p - 1]):
                assert ni[i, p] == rows[p], (i, p)
