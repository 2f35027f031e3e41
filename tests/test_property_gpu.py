"""Property test (hypothesis): random small tables built from adversarial ingredients -- duplicated
rows, zero rows, power-of-two rescaled rows, near-duplicates one ulp apart, random ids in the full
128-bit range -- random k and batch size, both scan paths, interleaved deletes.  The CUDA path must
equal the oracle bit for bit on every example."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import cosine_topk as O

pytestmark = pytest.mark.gpu
DIM = 1024


@st.composite
def tables(draw):
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    n_base = draw(st.integers(1, 40))
    base = rng.standard_normal((n_base, DIM)).astype(np.float32)
    rows = [base]
    if draw(st.booleans()):                                   # exact duplicates
        rows.append(base[rng.integers(0, n_base, size=draw(st.integers(1, 80)))])
    if draw(st.booleans()):                                   # power-of-two rescalings (same cosine)
        pick = base[rng.integers(0, n_base, size=draw(st.integers(1, 40)))]
        rows.append(pick * (2.0 ** rng.integers(-30, 31, size=(pick.shape[0], 1))).astype(np.float32))
    if draw(st.booleans()):                                   # one-ulp neighbours: near-ties below fp32 resolution
        pick = base[rng.integers(0, n_base, size=draw(st.integers(1, 40)))].copy()
        j = rng.integers(0, DIM, size=pick.shape[0])
        pick[np.arange(pick.shape[0]), j] = np.nextafter(pick[np.arange(pick.shape[0]), j], np.float32(np.inf))
        rows.append(pick)
    if draw(st.booleans()):                                   # zero-norm rows (NaN distance, sorted last)
        rows.append(np.zeros((draw(st.integers(1, 5)), DIM), np.float32))
    if draw(st.booleans()):                                   # bulk so that the tcgen05 path is eligible
        rows.append(rng.standard_normal((4200, DIM)).astype(np.float32))
    X = np.concatenate(rows)
    perm = rng.permutation(X.shape[0])
    X = X[perm]
    wide = draw(st.booleans())
    vals = rng.choice(2**40, size=X.shape[0], replace=False).astype(object)
    if wide:
        vals = [int(v) << 70 | int(rng.integers(0, 2**60)) for v in vals]
    ids = O.ids_from_ints([int(v) for v in vals])
    k = draw(st.sampled_from([1, 5, 12, 16, 17, 32, 40, 64, 65, 128]))
    nq = draw(st.sampled_from([1, 2, 7, 33]))
    qsrc = draw(st.sampled_from(["rows", "random", "mixed"]))
    if qsrc == "rows":
        Q = X[rng.integers(0, X.shape[0], size=nq)] * np.float32(draw(st.sampled_from([1.0, 0.25, 8.0])))
    elif qsrc == "random":
        Q = rng.standard_normal((nq, DIM)).astype(np.float32)
    else:
        Q = (X[rng.integers(0, X.shape[0], size=nq)] + 0.3 * rng.standard_normal((nq, DIM))).astype(np.float32)
    n_del = draw(st.integers(0, min(20, X.shape[0] - 1)))
    return X, ids, np.ascontiguousarray(Q), k, n_del, draw(st.sampled_from(["fp32", "bf16"]))


@settings(max_examples=80, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tables())
def test_engine_equals_oracle_on_adversarial_tables(case):
    import outline_rag_b200 as orx
    from tests._helpers import stored_bf16_rows
    X, ids, Q, k, n_del, dtype = case
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    with orx.Index(dtype) as ix:
        ix.upsert(ids, X)
        keep = np.ones(X.shape[0], bool)
        if n_del:
            keep[:n_del] = False
            assert ix.delete(ids[:n_del]) == n_del
        g_ids, g_d, g_c = ix.search(Q, k)
    for i in range(Q.shape[0]):
        w_ids, w_d = O.topk_exact(rows[keep], ids[keep], Q[i], k, exhaustive=rows.shape[0] <= 400)
        m = len(w_d)
        assert g_c[i] == m
        assert np.array_equal(g_ids[i, :m], w_ids), (dtype, k, i)
        assert np.array_equal(g_d[i, :m].view(np.uint64), w_d.view(np.uint64)), (dtype, k, i)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tables(), st.sampled_from([0.02, 0.5, 0.97]), st.integers(0, 2**31 - 1))
def test_filtered_search_equals_oracle_on_adversarial_tables(case, frac, seed):
    """The same tables under a random allow-list: the list regime (few eligible rows), the bitmap scan (single queries)
    and the masked-scale tensor-core pass (batches) must all equal the oracle on the eligible subset."""
    import outline_rag_b200 as orx
    from tests._helpers import stored_bf16_rows
    X, ids, Q, k, n_del, dtype = case
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    rng = np.random.default_rng(seed)
    with orx.Index(dtype) as ix:
        ix.upsert(ids, X)
        keep = np.ones(X.shape[0], bool)
        if n_del:
            keep[:n_del] = False
            ix.delete(ids[:n_del])
        allowed = rng.random(X.shape[0]) < frac
        allowed[rng.integers(0, X.shape[0])] = True
        unknown = O.ids_from_ints([2**100 + 7, 2**100 + 8])              # ids the table has never seen are ignored
        g_ids, g_d, g_c = ix.search_filtered(Q, k, np.concatenate([ids[allowed], unknown]))
    sel = keep & allowed
    for i in range(Q.shape[0]):
        if not sel.any():
            assert g_c[i] == 0
            continue
        w_ids, w_d = O.topk_exact(rows[sel], ids[sel], Q[i], k, exhaustive=True)
        m = len(w_d)
        assert g_c[i] == m
        assert np.array_equal(g_ids[i, :m], w_ids), (dtype, k, i, frac)
        nan = np.isnan(w_d)
        assert np.array_equal(np.isnan(g_d[i, :m]), nan)
        assert np.array_equal(g_d[i, :m][~nan].view(np.uint64), w_d[~nan].view(np.uint64)), (dtype, k, i, frac)
