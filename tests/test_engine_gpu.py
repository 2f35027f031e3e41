"""Parity tests proper: the CUDA path through the C-ABI vs the oracle, same seeded inputs.

Bar (BASELINE.json north_star): fp32 mode -> ids bit-exact (ties by chunk id); the canonical
distance is DEFINED as the fixed-order binary64 evaluation (oracle/cosine_topk.py:canon_distance)
and the CUDA rescore evaluates the identical tree, so distances are compared BIT FOR BIT too
(tolerance 0, far inside the 1e-5 relative the north star allows).  bf16 mode -> bit-exact
against the oracle run on the rows as stored (RNE bf16 of the normalised row) + recall@12 vs fp32.
"""
import json
import os

import numpy as np
import pytest

from oracle import cosine_topk as O
from tests._helpers import stored_bf16_rows

pytestmark = pytest.mark.gpu
DIM, K = 1024, 12


def _ids(n, start=0):
    return O.ids_arange(start, start + n)


def _check_exact(index, X, ids, Q, k=K, oracle_rows=None):
    Xo = X if oracle_rows is None else oracle_rows
    g_ids, g_d, g_c = index.search(Q, k)
    for i in range(Q.shape[0]):
        w_ids, w_d = O.topk_exact(Xo, ids, Q[i], k)
        m = len(w_d)
        assert g_c[i] == m
        assert np.array_equal(g_ids[i, :m], w_ids), f"query {i}: ids differ"
        assert np.array_equal(g_d[i, :m].view(np.uint64), w_d.view(np.uint64)), f"query {i}: distance bits differ"
        assert np.isnan(g_d[i, m:]).all() and (g_ids[i, m:] == 0).all()
    return g_ids, g_d, g_c


@pytest.fixture()
def Index():
    import outline_rag_b200 as orx
    return orx.Index


# ------------------------------------------------------------------ seeded parity, both dtypes
def test_fp32_bit_exact_vs_oracle(Index, small_table):
    X, Q, anchors = small_table
    ids = _ids(X.shape[0])
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        assert len(ix) == X.shape[0]
        g_ids, _, _ = _check_exact(ix, X, ids, Q)
        assert (g_ids[:, 0, 1] == anchors.astype(np.uint64)).all()
        st = ix.stats()
        assert st["kernel_launches"] > 0 and st["last_path"] in (1, 2)


def test_bf16_bit_exact_vs_oracle_on_stored_rows_and_recall(Index, small_table):
    X, Q, _ = small_table
    ids = _ids(X.shape[0])
    Xb = stored_bf16_rows(X)
    with Index("bf16") as ix:
        ix.upsert(ids, X)
        got, _ = ix.fetch(ids[:64])
        assert np.array_equal(got.view(np.uint32), Xb[:64].view(np.uint32))
        g_ids, _, _ = _check_exact(ix, X, ids, Q, oracle_rows=Xb)
    rec = np.mean([O.recall_at_k(g_ids[i], O.topk_exact(X, ids, Q[i], K)[0]) for i in range(Q.shape[0])])
    assert rec >= 0.95, rec


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_golden_vectors_through_the_cuda_path(Index, synth100k, dtype):
    with open(os.path.join(os.path.dirname(__file__), "golden", "topk_golden.json")) as f:
        G = json.load(f)
    X = synth100k.table(G["n_rows"])
    Q, _ = synth100k.queries(len(G["cases"]), G["n_rows"])
    with Index(dtype) as ix:
        ix.upsert(_ids(G["n_rows"]), X)
        g_ids, g_d, g_c = ix.search(Q, G["k"])
    pre = "" if dtype == "fp32" else "bf16_"
    for qi, case in enumerate(G["cases"]):
        assert O.ids_to_ints(g_ids[qi]) == case[pre + "ids"]
        assert [float(d).hex() for d in g_d[qi]] == case[pre + "dist_hex"]


@pytest.mark.parametrize("nq", [1, 2, 5, 33, 64, 130])
def test_query_batches(Index, small_table, nq):
    X, Q, _ = small_table
    ids = _ids(4096)
    rng = np.random.default_rng(nq)
    Qb = np.concatenate([Q, rng.standard_normal((max(0, nq - Q.shape[0]), DIM)).astype(np.float32)])[:nq]
    with Index("fp32") as ix:
        ix.upsert(ids, X[:4096])
        _check_exact(ix, X[:4096], ids, Qb)


@pytest.mark.parametrize("k", [1, 2, 12, 16, 17, 32, 33, 64, 96, 100, 128])
def test_all_k(Index, small_table, k):
    X, Q, _ = small_table
    ids = _ids(3000)
    with Index("fp32") as ix:
        ix.upsert(ids, X[:3000])
        _check_exact(ix, X[:3000], ids, Q[:6], k=k)


@pytest.mark.parametrize("n", [1, 2, 11, 12, 13, 31, 32, 33, 63, 64, 65, 255, 1023, 2049])
def test_ragged_table_sizes(Index, small_table, n):
    X, Q, _ = small_table
    with Index("fp32") as ix:
        ix.upsert(_ids(n), X[:n])
        _check_exact(ix, X[:n], _ids(n), Q[:3])


# ------------------------------------------------------------------ known answers
def test_ka1_identity_basis(Index):
    X = np.eye(DIM, dtype=np.float32)
    q = np.arange(DIM, 0, -1).astype(np.float32)
    with Index("fp32") as ix:
        ix.upsert(_ids(DIM), X)
        g_ids, g_d, _ = ix.search(q, K)
    assert O.ids_to_ints(g_ids[0]) == list(range(K))
    want = 1.0 - (DIM - np.arange(K)) / np.sqrt(np.sum(q.astype(np.float64) ** 2))
    np.testing.assert_allclose(g_d[0], want, rtol=0, atol=1e-15)


def test_ka2_duplicate_rows_order_by_id(Index):
    rng = np.random.default_rng(1)
    base = rng.standard_normal((40, DIM)).astype(np.float32)
    X = np.concatenate([base] * 5)                       # 200 rows, every vector five times
    idv = rng.permutation(200) + 1000
    ids = O.ids_from_ints(idv.tolist())
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        _check_exact(ix, X, ids, base[:4], k=12)
        g_ids, g_d, _ = ix.search(base[7], 5)
    same = sorted(int(idv[7 + 40 * j]) for j in range(5))
    assert O.ids_to_ints(g_ids[0]) == same and (g_d[0] == g_d[0][0]).all()


def test_ka2b_many_exact_ties_force_the_exhaustive_path(Index):
    """More identical rows than the candidate list holds: completeness cannot be proven from
    the list alone, so the threshold-collect fallback must produce the id-ascending answer."""
    rng = np.random.default_rng(11)
    v = rng.standard_normal(DIM).astype(np.float32)
    X = np.concatenate([np.tile(v, (500, 1)), rng.standard_normal((1500, DIM)).astype(np.float32)])
    idv = rng.permutation(2000) + 7
    ids = O.ids_from_ints(idv.tolist())
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        g_ids, g_d, g_c = ix.search(v, K)
        st = ix.stats()
    assert O.ids_to_ints(g_ids[0]) == sorted(idv[:500].tolist())[:K]
    assert (g_d[0] == 0.0).all() or np.allclose(g_d[0], 0.0, atol=1e-15)
    assert st["fallback_exhaustive"] >= 1


def test_ka3_row_scale_invariance(Index, small_table):
    X, Q, _ = small_table
    X = X[:2000]
    rng = np.random.default_rng(2)
    s = (2.0 ** rng.integers(-20, 21, size=2000)).astype(np.float32)
    ids = _ids(2000)
    with Index("fp32") as a, Index("fp32") as b:
        a.upsert(ids, X)
        b.upsert(ids, X * s[:, None])
        ra, rb = a.search(Q[:8], K), b.search(Q[:8], K)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1].view(np.uint64), rb[1].view(np.uint64))


def test_ka3b_extreme_magnitudes_are_still_exact(Index, small_table):
    """Rows whose norm is far outside fp32's comfortable range are never pruned by the fast scan."""
    X, Q, _ = small_table
    X = X[:1500].copy()
    X[5] *= np.float32(1e-30)
    X[6] *= np.float32(1e30)
    X[7] *= np.float32(1e-44)       # denormal elements
    ids = _ids(1500)
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        _check_exact(ix, X, ids, np.stack([X[5], X[6], Q[0], Q[1]]) / np.float32(1.0))


def test_ka4_antiparallel_row_distance_two(Index):
    rng = np.random.default_rng(3)
    q = rng.standard_normal(DIM).astype(np.float32)
    X = np.stack([q, -q, 4 * q])
    with Index("fp32") as ix:
        ix.upsert(_ids(3), X)
        g_ids, g_d, g_c = ix.search(q, K)
    assert g_c[0] == 3 and O.ids_to_ints(g_ids[0, :3]) == [0, 2, 1]
    assert g_d[0, 0] == 0.0 and g_d[0, 1] == 0.0 and g_d[0, 2] == 2.0


def test_ka5_zero_norm_rows_sort_last(Index, small_table):
    X, Q, _ = small_table
    X = X[:40].copy()
    X[[3, 17]] = 0.0
    ids = _ids(40)
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        g_ids, g_d, g_c = ix.search(Q[0], 32)
        assert g_c[0] == 32 and not np.isnan(g_d[0]).any()           # 38 finite rows >= 32
        ix.delete(ids[20:])
        _check_exact(ix, X[:20], ids[:20], Q[:2], k=32)              # 18 finite + 2 NaN rows, NaN last by id
        g_ids, g_d, g_c = ix.search(Q[0], 32)
    assert g_c[0] == 20 and np.isnan(g_d[0, 18:20]).all() and O.ids_to_ints(g_ids[0, 18:20]) == [3, 17]


def test_zero_query_gives_all_nan_by_id(Index, small_table):
    X, _, _ = small_table
    ids = _ids(100)
    with Index("fp32") as ix:
        ix.upsert(ids, X[:100])
        g_ids, g_d, g_c = ix.search(np.zeros(DIM, np.float32), K)
    assert g_c[0] == K and np.isnan(g_d[0]).all() and O.ids_to_ints(g_ids[0]) == list(range(K))


def test_ka6_k_larger_than_table_and_empty_table(Index, small_table):
    X, Q, _ = small_table
    with Index("fp32") as ix:
        g_ids, g_d, g_c = ix.search(Q[:2], K)
        assert (g_c == 0).all() and np.isnan(g_d).all()
        ix.upsert(_ids(7), X[:7])
        _check_exact(ix, X[:7], _ids(7), Q[:2])


def test_ka7_delete_and_upsert_semantics(Index, small_table):
    X, Q, _ = small_table
    n = 3000
    ids = _ids(n)
    with Index("fp32") as ix:
        ix.upsert(ids, X[:n])
        top = ix.search(Q[0], K)[0][0]
        assert ix.delete(top[:5]) == 5
        assert ix.delete(top[:5]) == 0                     # unknown ids are ignored, like SQL DELETE
        assert len(ix) == n - 5 and not ix.contains(int(top[0, 1]))
        keep = np.ones(n, bool)
        keep[top[:5, 1].astype(np.int64)] = False
        _check_exact(ix, X[:n][keep], ids[keep], Q[:4])
        # upsert of an existing id replaces the row in place; size unchanged
        victim = int(top[7, 1])
        ix.upsert([victim], X[n + 1][None])
        assert len(ix) == n - 5
        Xm = X[:n].copy()
        Xm[victim] = X[n + 1]
        _check_exact(ix, Xm[keep], ids[keep], np.stack([Q[0], X[n + 1]]))
        # the same id twice in one batch: last occurrence wins
        ix.upsert([victim, victim], np.stack([X[n + 2], X[n + 3]]))
        Xm[victim] = X[n + 3]
        _check_exact(ix, Xm[keep], ids[keep], np.stack([X[n + 2], X[n + 3]]))
        assert ix.stats()["rows_moved"] > 0


def test_delete_everything_then_refill(Index, small_table):
    X, Q, _ = small_table
    ids = _ids(500)
    with Index("fp32", capacity=16) as ix:          # also exercises table growth
        ix.upsert(ids, X[:500])
        assert ix.capacity >= 500
        assert ix.delete(ids) == 500 and len(ix) == 0
        assert ix.search(Q[0], K)[2][0] == 0
        ix.upsert(ids[:100], X[100:200])
        _check_exact(ix, X[100:200], ids[:100], Q[:2])


def test_ka8_input_errors(Index, small_table):
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    with Index("fp32") as ix:
        ix.upsert(_ids(10), X[:10])
        with pytest.raises(orx.OrxValueError, match="dimensions"):
            ix.upsert(_ids(2), np.zeros((2, 768), np.float32))
        bad = X[:3].copy()
        bad[1, 5] = np.nan
        with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
            ix.upsert(_ids(3, 100), bad)
        assert len(ix) == 10                                  # whole batch rejected, table unchanged
        bad[1, 5] = np.inf
        with pytest.raises(orx.OrxValueError):
            ix.upsert(_ids(3, 100), bad)
        with pytest.raises(orx.OrxValueError, match="dimensions"):
            ix.search(np.zeros((1, 512), np.float32), K)
        qbad = Q[:2].copy()
        qbad[1, 0] = np.nan
        with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
            ix.search(qbad, K)
        for k in (0, 129, -1):
            with pytest.raises(orx.OrxValueError):
                ix.search(Q[:1], k)
        _check_exact(ix, X[:10], _ids(10), Q[:2])             # still healthy afterwards


def test_128_bit_ids_and_uuid_order(Index, small_table):
    import uuid
    X, Q, _ = small_table
    rng = np.random.default_rng(9)
    vals = [int(rng.integers(0, 2**63)) << 64 | int(rng.integers(0, 2**63)) for _ in range(300)]
    vals[10] = (1 << 127) | 5                                  # top bit set: unsigned order matters
    ids = O.ids_from_ints(vals)
    Xd = np.concatenate([X[:150], X[:150]])                    # duplicates across different high words
    with Index("fp32") as ix:
        ix.upsert([str(uuid.UUID(int=v)) for v in vals], Xd)
        _check_exact(ix, Xd, ids, Q[:4])
        assert ix.contains(str(uuid.UUID(int=vals[10])))


# ------------------------------------------------------------------ device hand-off + merge
def test_device_pointers_in_and_out(Index, small_table):
    import torch
    X, Q, _ = small_table
    ids = _ids(2048)
    with Index("fp32") as ix:
        ix.use_torch_stream()
        ix.upsert(ids, torch.from_numpy(X[:2048]).cuda())
        h = ix.search(Q[:5], K)
        d = ix.search(torch.from_numpy(Q[:5]).cuda(), K)
        assert d[0].is_cuda
        assert np.array_equal(d[0].cpu().numpy().view(np.uint64), h[0])
        assert np.array_equal(d[1].cpu().numpy().view(np.uint64), h[1].view(np.uint64))
        assert np.array_equal(d[2].cpu().numpy(), h[2])


@pytest.mark.parametrize("n_shards", [2, 4, 8])
def test_shard_merge_equals_single_table(Index, small_table, n_shards):
    """The multi-GPU path emulated on one GPU: G row shards (same partition function as
    sharded.py), G local searches, orx_merge_topk == the single-table answer, bit for bit."""
    import torch
    from outline_rag_b200.sharded import shard_of
    X, Q, _ = small_table
    n = 6000
    ids = _ids(n)
    owner = shard_of(ids, n_shards)
    shards = [Index("fp32") for _ in range(n_shards)]
    try:
        for s, ix in enumerate(shards):
            ix.upsert(ids[owner == s], X[:n][owner == s])
        parts = [ix.search(Q[:9], K) for ix in shards]
        g_ids = np.stack([p[0] for p in parts])
        g_d = np.stack([p[1] for p in parts])
        g_c = np.stack([p[2] for p in parts])
        m_ids, m_d, m_c = shards[0].merge_topk(g_ids, g_d, g_c, K)
        dm = shards[0].merge_topk(torch.from_numpy(g_ids.view(np.int64)).cuda(), torch.from_numpy(g_d).cuda(),
                                  torch.from_numpy(g_c).cuda(), K)
    finally:
        for ix in shards:
            ix.close()
    for i in range(9):
        w_ids, w_d = O.topk_exact(X[:n], ids, Q[i], K)
        assert np.array_equal(m_ids[i], w_ids) and np.array_equal(m_d[i].view(np.uint64), w_d.view(np.uint64))
    assert np.array_equal(dm[0].cpu().numpy().view(np.uint64), m_ids)
    assert (m_c == K).all()


def test_merge_handles_short_lists(Index, small_table):
    X, Q, _ = small_table
    with Index("fp32") as a, Index("fp32") as b:
        a.upsert(_ids(5), X[:5])
        b.upsert(_ids(3, 100), X[5:8])
        pa, pb = a.search(Q[:2], K), b.search(Q[:2], K)
        m = a.merge_topk(np.stack([pa[0], pb[0]]), np.stack([pa[1], pb[1]]), np.stack([pa[2], pb[2]]), K)
    allids = np.concatenate([_ids(5), _ids(3, 100)])
    for i in range(2):
        w_ids, w_d = O.topk_exact(X[:8], allids, Q[i], K)
        assert m[2][i] == 8 and np.array_equal(m[0][i, :8], w_ids)
        assert np.array_equal(m[1][i, :8].view(np.uint64), w_d.view(np.uint64))


# ------------------------------------------------------------------ tcgen05 batched scan
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("nq", [2, 32, 128, 129, 300])
def test_tcgen05_batched_scan_is_exact(Index, synth100k, dtype, nq):
    """Query batches run the TMA + tcgen05 coarse scan with the fused threshold-collect epilogue;
    results must still equal the oracle bit for bit (canonical rescore + completeness proof)."""
    n = 20000
    X = synth100k.table(n)
    Q, _ = synth100k.queries(nq, n)
    ids = _ids(n)
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    with Index(dtype) as ix:
        ix.upsert(ids, X)
        _check_exact(ix, X, ids, Q, oracle_rows=rows)
        st = ix.stats()
    assert st["last_path"] == 2
    assert st["fallback_gemv"] <= max(1, nq // 16), st        # the coarse pass proves almost every query itself


def test_tcgen05_ragged_rows_special_rows_and_k32(Index, synth100k):
    n = 4096 + 77                                             # last tile is partial
    X = synth100k.table(n).copy()
    X[100] = 0.0                                              # zero-norm row: NaN distance, never a candidate
    X[200] *= np.float32(1e30)                                # irregular magnitude: always a candidate
    X[4100] *= np.float32(1e-30)
    Q, _ = synth100k.queries(40, n)
    Q = np.concatenate([Q, X[200:201] / np.float32(1e30), X[4100:4101] * np.float32(1e30)])
    ids = _ids(n)
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        _check_exact(ix, X, ids, Q, k=32)
        _check_exact(ix, X, ids, Q, k=1)
        assert ix.stats()["last_path"] == 2


def test_tcgen05_duplicates_overflow_falls_back(Index, synth100k):
    n = 6000
    X = synth100k.table(n).copy()
    X[1000:1200] = X[5]                                       # 200 identical rows: wider than any candidate list
    rng = np.random.default_rng(3)
    idv = rng.permutation(n) + 10
    ids = O.ids_from_ints(idv.tolist())
    Q = np.stack([X[5], X[7], X[9]])
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        _check_exact(ix, X, ids, Q)
        st = ix.stats()
    assert st["fallback_gemv"] >= 1 and st["fallback_exhaustive"] >= 1


def test_gathered_block_merge_equals_single_table(Index, small_table):
    """The NCCL path's data flow on one GPU: every shard searches INTO views of its result block,
    the blocks are concatenated the way all_gather_into_tensor leaves them, and
    orx_merge_topk_strided reads the gathered buffer in place."""
    import torch
    from outline_rag_b200.sharded import ShardedIndex, shard_of
    X, Q, _ = small_table
    n, G, nq = 6000, 4, 7
    ids = _ids(n)
    owner = shard_of(ids, G)
    qd = torch.from_numpy(Q[:nq]).cuda()
    shards = [Index("fp32") for _ in range(G)]
    try:
        host = ShardedIndex(local_index=shards[0])
        host.world = G                                   # lay the plan out for G ranks
        p = host._plan(nq, K, qd.device)
        blocks = []
        for s, ix in enumerate(shards):
            ix.upsert(ids[owner == s], X[:n][owner == s])
            ix.search_into(qd, K, p["ids"], p["dist"], p["cnt"])
            blocks.append(p["block"].clone())
        p["gathered"].copy_(torch.cat(blocks))
        shards[0].merge_blocks(p["gathered"], G, nq, K, p["words"] * 8, p["out_ids"], p["out_dist"], p["out_cnt"])
        torch.cuda.synchronize()
        m_ids = p["out_ids"].cpu().numpy().view(np.uint64)
        m_d = p["out_dist"].cpu().numpy()
        assert (p["out_cnt"].cpu().numpy() == K).all()
    finally:
        for ix in shards:
            ix.close()
    for i in range(nq):
        w_ids, w_d = O.topk_exact(X[:n], ids, Q[i], K)
        assert np.array_equal(m_ids[i], w_ids) and np.array_equal(m_d[i].view(np.uint64), w_d.view(np.uint64))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_snapshot_roundtrip_is_bit_identical(Index, small_table, tmp_path, dtype):
    """save -> load reproduces the table verbatim (bf16 rows are NOT re-normalised), so searches and
    fetches are bit-identical; a snapshot with a NaN row is rejected without side effects."""
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    n = 5000
    ids = _ids(n, 7)
    with Index(dtype) as a:
        a.upsert(ids, X[:n])
        a.delete(ids[10:30])                                    # compaction reorders rows: the snapshot keeps that order
        man = a.save(str(tmp_path / "snap"))
        assert man["rows"] == n - 20 and man["dtype"] == dtype
        ra = a.search(Q[:6], K)
        fa, _ = a.fetch(ids[:64])
    b = orx.Index.load(str(tmp_path / "snap"))
    try:
        assert len(b) == n - 20 and b.dtype == dtype
        rb = b.search(Q[:6], K)
        fb, found = b.fetch(ids[:64])
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1].view(np.uint64), rb[1].view(np.uint64))
        assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32)) and found[:10].all() and not found[10:30].any()
        b.upsert(ids[10:12], X[10:12])                          # the restored table keeps working
        assert len(b) == n - 18
    finally:
        b.close()
    # corrupt the snapshot with a NaN: load must fail
    v = np.load(tmp_path / "snap" / "vecs.npy", mmap_mode="r+")
    v[3, 0:4 if dtype == "fp32" else 2] = 0xFF
    v.flush()
    with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
        orx.Index.load(str(tmp_path / "snap"))


def test_snapshot_is_atomic_and_refuses_a_table_that_changes_under_it(Index, small_table, tmp_path, monkeypatch):
    """`save` exported in chunks with the lock released in between and overwrote the files in place.  Now
    it writes beside the target and renames, and it notices a write that lands between two chunks (mutation counter)."""
    import os
    import outline_rag_b200 as orx
    from outline_rag_b200 import engine as E
    X, Q, _ = small_table
    ids = _ids(3000)
    snap = str(tmp_path / "snap")
    with Index("fp32") as a:
        a.upsert(ids, X[:3000])
        a.save(snap)
        a.delete(ids[:100])
        man = a.save(snap)                                      # replaces the older snapshot in one rename
        assert man["rows"] == 2900 and sorted(os.listdir(tmp_path)) == ["snap"]
        with orx.Index.load(snap) as b:
            assert len(b) == 2900 and not b.contains(5)
        # a writer slips in between two export chunks: the half-taken snapshot is discarded, the old one survives
        monkeypatch.setattr(E.Index, "SNAPSHOT_CHUNK", 1000)
        real = E.lib.orx_export_rows
        calls = {"n": 0}

        def export_and_write(*args):
            calls["n"] += 1
            rc = real(*args)
            if calls["n"] % 2 == 1:
                a.upsert(_ids(1, 900_000 + calls["n"]), X[4000:4001])       # every attempt sees a new write
            return rc

        monkeypatch.setattr(E.lib, "orx_export_rows", export_and_write)
        with pytest.raises(orx.OrxError, match="kept changing"):
            a.save(snap, retries=2)
        monkeypatch.setattr(E.lib, "orx_export_rows", real)
        assert sorted(os.listdir(tmp_path)) == ["snap"]
        with orx.Index.load(snap) as b:
            assert len(b) == 2900                               # still the last GOOD snapshot
        assert a.save(snap)["rows"] == len(a)


def test_concurrent_threads_search_while_a_writer_refreshes(Index, small_table):
    """The uvicorn worker runs searches and the refresh task interleaved (asyncio.to_thread): calls on one
    index are serialised, every search sees a consistent table (either before or after a whole batch)."""
    import threading
    X, Q, _ = small_table
    n = 4000
    ids = _ids(n)
    with Index("fp32") as ix:
        ix.upsert(ids, X[:n])
        want_before = [O.topk_exact(X[:n], ids, Q[i], K)[0] for i in range(4)]
        Xa = X[:n].copy()
        Xa[:200] = X[n:n + 200]
        want_after = [O.topk_exact(Xa, ids, Q[i], K)[0] for i in range(4)]
        errors, seen = [], {"before": 0, "after": 0}

        def reader(i):
            try:
                for _ in range(30):
                    got = ix.search(Q[i], K)[0][0]
                    if np.array_equal(got, want_before[i]):
                        seen["before"] += 1
                    elif np.array_equal(got, want_after[i]):
                        seen["after"] += 1
                    else:
                        errors.append("inconsistent result")
            except Exception as e:      # noqa: BLE001
                errors.append(repr(e))

        def writer():
            try:
                ix.upsert(ids[:200], X[n:n + 200])         # in-place replacement of 200 rows, one batch
            except Exception as e:      # noqa: BLE001
                errors.append(repr(e))

        threads = [threading.Thread(target=reader, args=(i,)) for i in range(4)] + [threading.Thread(target=writer)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors[:3]
        assert seen["before"] + seen["after"] == 120
        got = ix.search(Q[:4], K)[0]
        assert all(np.array_equal(got[i], want_after[i]) for i in range(4))


def test_more_queries_than_one_tcgen05_launch_takes(Index, synth100k):
    """nq > 2048 is split into several tcgen05 launches (and > 1024 into several exchange rounds)."""
    n, nq = 5000, 2100
    X = synth100k.table(n)
    rng = np.random.default_rng(5)
    Q = (X[rng.integers(0, n, size=nq)] + 0.2 * rng.standard_normal((nq, DIM))).astype(np.float32)
    ids = _ids(n)
    with Index("bf16") as ix:
        ix.upsert(ids, X)
        g_ids, g_d, g_c = ix.search(Q, K)
        assert ix.stats()["last_path"] == 2
        ix.shard_connect([ix.shard_export(1, 0)])
        s_ids, s_d, s_c = ix.search_sharded(Q, K)
    assert np.array_equal(g_ids, s_ids) and np.array_equal(g_d.view(np.uint64), s_d.view(np.uint64))
    rows = stored_bf16_rows(X)
    for i in list(range(0, nq, 97)) + [2047, 2048, 2099]:
        w_ids, w_d = O.topk_exact(rows, ids, Q[i], K)
        assert np.array_equal(g_ids[i], w_ids) and np.array_equal(g_d[i].view(np.uint64), w_d.view(np.uint64)), i
    assert (g_c == K).all()


def test_host_upsert_larger_than_one_staging_chunk(Index, synth100k):
    """70 000 host rows = two 65 536-row staging chunks; a NaN in the SECOND chunk rejects the whole batch."""
    import outline_rag_b200 as orx
    n = 70_000
    X = synth100k.table(n)
    ids = _ids(n)
    Q, _ = synth100k.queries(3, n)
    with Index("fp32", capacity=1024) as ix:
        bad = X.copy()
        bad[69_000, 17] = np.nan
        with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
            ix.upsert(ids, bad)
        assert len(ix) == 0
        ix.upsert(ids, X)
        assert len(ix) == n
        got, found = ix.fetch(ids[[0, 65_535, 65_536, 69_999]])
        assert found.all() and np.array_equal(got.view(np.uint32), X[[0, 65_535, 65_536, 69_999]].view(np.uint32))
        _check_exact(ix, X, ids, Q)


def test_filtered_search_is_the_sql_with_a_where_clause(Index, small_table):
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    n = 3000
    X = X[:n].copy()
    X[77] = 0.0                                                # an eligible zero-norm row sorts last
    ids = _ids(n, 500)
    rng = np.random.default_rng(8)
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        for m in (0, 1, 5, 12, 40, 700):
            sel = np.sort(rng.choice(n, size=m, replace=False)) if m else np.zeros(0, np.int64)
            if m >= 5:
                sel[0] = 77
                sel = np.unique(sel)
            allow = np.concatenate([ids[sel], ids[sel][:3], O.ids_from_ints([7, 8])]) if m else O.ids_from_ints([7])
            g_ids, g_d, g_c = ix.search_filtered(Q[:3], K, allow)          # duplicates + unknown ids are harmless
            for i in range(3):
                w_ids, w_d = O.topk_exact(X[sel], ids[sel], Q[i], K, exhaustive=True)
                mm = len(w_d)
                assert g_c[i] == mm
                assert np.array_equal(g_ids[i, :mm], w_ids)
                assert np.array_equal(g_d[i, :mm].view(np.uint64), w_d.view(np.uint64))
                assert np.isnan(g_d[i, mm:]).all()
        bad = Q[:1].copy()
        bad[0, 0] = np.nan
        with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
            ix.search_filtered(bad, K, ids[:5])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_filtered_search_with_many_eligible_rows_scans_through_a_row_bitmap(Index, small_table, dtype):
    """>= 4096 eligible rows: the predicate becomes a bitmap and the sequential scan skips the other rows
    (csrc/scan_gemv.cu scan_gemv_filtered_kernel).  Same answers as the oracle on the eligible subset."""
    X, Q, _ = small_table                                      # 8192 rows
    n = X.shape[0]
    X = X.copy()
    X[4000] = 0.0                                              # eligible zero-norm row
    X[100:140] = X[99]                                         # 41 identical eligible rows: a tie wider than the list
    ids = _ids(n, 1000)
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    rng = np.random.default_rng(21)
    with Index(dtype) as ix:
        ix.upsert(ids, X)
        for m in (4096, 5000, n):
            sel = np.sort(rng.choice(n, size=m, replace=False))
            if m < n:
                sel = np.unique(np.concatenate([sel, np.arange(99, 140), [4000]]))
            queries = np.concatenate([Q[:4], X[99:100], np.zeros((1, DIM), np.float32)])   # a tie query, a zero query
            for k in (K, 40):
                launches0 = ix.stats()["scan_launches"]
                g_ids, g_d, g_c = ix.search_filtered(queries, k, ids[rng.permutation(sel)])
                assert ix.stats()["scan_launches"] > launches0                              # the scan ran
                for i in range(queries.shape[0]):
                    w_ids, w_d = O.topk_exact(rows[sel], ids[sel], queries[i], k, exhaustive=True)
                    assert g_c[i] == k
                    assert np.array_equal(g_ids[i], w_ids), (m, k, i)
                    nan = np.isnan(w_d)
                    assert np.array_equal(np.isnan(g_d[i]), nan), (m, k, i)
                    assert np.array_equal(g_d[i][~nan].view(np.uint64), w_d[~nan].view(np.uint64)), (m, k, i)
        # every row eligible == the unfiltered search
        a = ix.search(Q[:8], K)
        b = ix.search_filtered(Q[:8], K, ids)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))
        # the bitmap follows deletes (rows move) because it is rebuilt from the ids on every call
        gone = ids[sel[:50]]
        ix.delete(gone)
        keep = np.ones(n, bool)
        keep[sel[:50]] = False
        sel2 = np.array([r for r in sel if keep[r]])
        g_ids, g_d, _ = ix.search_filtered(Q[:2], K, ids[sel])                              # deleted ids are unknown now
        for i in range(2):
            w_ids, w_d = O.topk_exact(rows[sel2], ids[sel2], Q[i], K, exhaustive=True)
            assert np.array_equal(g_ids[i], w_ids) and np.array_equal(g_d[i].view(np.uint64), w_d.view(np.uint64))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_filtered_batches_take_one_tensor_core_pass(Index, synth100k, dtype):
    """A batch of filtered queries whose bitmap scans would together read more than the table runs ONE tcgen05 pass: the
    predicate is folded into the per-row scale (a cleared bit = the never-a-candidate marker).  Same answers as the oracle
    on the eligible subset -- ids and distance bits -- for one query tile and for CTA pairs; batches too small for the
    pass to pay keep the bitmap GEMV scan."""
    n = 20_000
    X = synth100k.table(n).copy()
    Q, _ = synth100k.queries(200, n)
    X[4000] = 0.0                                              # eligible zero-norm row
    X[4001] = 0.0                                              # excluded zero-norm row
    X[100:140] = X[99]                                         # 41 identical eligible rows
    X[7000:7100] = X[6999]                                     # 101 identical rows, NOT eligible: must not flood anything
    X[300] *= 1e30                                             # irregular magnitude ("always a candidate"), eligible
    X[301] *= 1e30                                             # the same, excluded
    ids = _ids(n, 1000)
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    rng = np.random.default_rng(5)
    excluded = np.concatenate([[4001, 301], np.arange(6999, 7100)])
    with Index(dtype) as ix:
        ix.upsert(ids, X)
        for m in (4500, 15_000):
            sel = rng.choice(n, size=m, replace=False)
            sel = np.setdiff1d(np.union1d(sel, np.concatenate([np.arange(99, 140), [4000, 300]])), excluded)
            queries = np.concatenate([Q, X[99:100], X[6999:7000], np.zeros((1, DIM), np.float32)])
            for nq in (9, len(queries)):
                qs = queries[-nq:]
                g_ids, g_d, g_c = ix.search_filtered(qs, K, ids[rng.permutation(sel)])
                assert ix.stats()["last_path"] == 2, (m, nq)
                # the three special queries at the end always; of a long batch otherwise every 6th (the oracle's
                # exhaustive pass is what takes the time here)
                for i in sorted(set(range(0, nq, 1 if nq < 20 else 6)) | set(range(nq - 3, nq))):
                    w_ids, w_d = O.topk_exact(rows[sel], ids[sel], qs[i], K, exhaustive=True)
                    assert g_c[i] == K
                    assert np.array_equal(g_ids[i], w_ids), (m, nq, i)
                    nan = np.isnan(w_d)
                    assert np.array_equal(np.isnan(g_d[i]), nan), (m, nq, i)
                    assert np.array_equal(g_d[i][~nan].view(np.uint64), w_d[~nan].view(np.uint64)), (m, nq, i)
            # the same filter: a short batch (bitmap GEMV scan unless 2 x eligible >= rows), a wide k
            a = ix.search_filtered(queries[:2], K, ids[sel])
            assert ix.stats()["last_path"] == (2 if 2 * len(sel) >= n else 1)
            b = ix.search_filtered(queries[:9], K, ids[sel])
            assert ix.stats()["last_path"] == (2 if 9 * len(sel) >= n else 1)
            assert np.array_equal(a[0], b[0][:2]) and np.array_equal(a[1].view(np.uint64), b[1][:2].view(np.uint64))
            w = ix.search_filtered(queries[:9], 40, ids[sel])                   # k > 32: the 160-key candidate lists
            assert ix.stats()["last_path"] == (2 if 9 * len(sel) >= n else 1)
            for i in (0, 8):
                w_ids, w_d = O.topk_exact(rows[sel], ids[sel], queries[i], 40, exhaustive=True)
                assert np.array_equal(w[0][i], w_ids) and np.array_equal(w[1][i].view(np.uint64), w_d.view(np.uint64))
        # a filter handle resolved once serves batches through the same pass
        with ix.make_filter(ids[sel]) as f:
            c = ix.search_filtered(queries[:9], K, f)
            assert ix.stats()["last_path"] == 2
            assert np.array_equal(c[0], b[0]) and np.array_equal(c[1].view(np.uint64), b[1].view(np.uint64))
        # fewer eligible regular rows than k in reach of the coarse pass: the exact list path answers
        few = np.concatenate([np.arange(5000, 5000 + 4096)])
        g = ix.search_filtered(queries[:6], K, ids[few])
        for i in range(6):
            w_ids, w_d = O.topk_exact(rows[few], ids[few], queries[i], K, exhaustive=True)
            assert np.array_equal(g[0][i], w_ids) and np.array_equal(g[1][i].view(np.uint64), w_d.view(np.uint64))


def test_filter_handle_is_resolved_once_and_follows_upserts_and_deletes(Index, small_table):
    """`orx_filter_*`: same answers as the per-call filter; the device bitmap is rebuilt only when the id -> row
    map has changed, and then denotes the live rows whose id is in the set."""
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    n = X.shape[0]
    ids = _ids(n, 10)
    rng = np.random.default_rng(4)
    sel = np.sort(rng.choice(n, size=5000, replace=False))
    few = sel[:30]
    with Index("fp32") as ix, Index("fp32") as other:
        ix.upsert(ids[:6000], X[:6000])
        with ix.make_filter(ids[sel]) as big, ix.make_filter(ids[few]) as small:
            def check(flt, rows_live):
                got = ix.search_filtered(Q[:3], K, flt)
                for i in range(3):
                    w_ids, w_d = O.topk_exact(X[rows_live], ids[rows_live], Q[i], K, exhaustive=True)
                    assert got[2][i] == len(w_d)
                    assert np.array_equal(got[0][i, :len(w_d)], w_ids)
                    assert np.array_equal(got[1][i, :len(w_d)].view(np.uint64), w_d.view(np.uint64))
            live = sel[sel < 6000]                                  # only part of the set is in the table yet
            assert len(live) < 4096                                 # ... so this is the list regime
            check(big, live)
            check(small, few[few < 6000])
            ix.upsert(ids[6000:], X[6000:])                         # the rest arrives: bitmap regime from now on
            launches0 = ix.stats()["scan_launches"]
            check(big, sel)
            check(big, sel)                                         # second use: nothing to resolve
            assert ix.stats()["scan_launches"] == launches0 + 2
            check(small, few)
            ix.upsert(ids[sel[:5]], X[sel[:5]])                     # overwrite in place: same rows, same answers
            check(big, sel)
            assert ix.delete(ids[sel[100:900]]) == 800              # rows move; 4200 eligible rows remain
            keep = np.ones(5000, bool)
            keep[100:900] = False
            check(big, sel[keep])
            assert ix.delete(ids[sel[900:1200]]) == 300             # 3900 left: back to the list regime
            keep[900:1200] = False
            check(big, sel[keep])
            with pytest.raises(orx.OrxValueError, match="another index"):
                other.search_filtered(Q[:1], K, big)
        with pytest.raises(orx.OrxValueError, match="closed"):
            ix.search_filtered(Q[:1], K, big)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_wide_k_batches_on_the_tensor_path(Index, synth100k, dtype):
    """16 < k <= 64 with a batch: one tcgen05 pass whose per-(query, CTA) candidate lists hold 160 keys (k <= 16: 64) --
    a list has to hold the rows within the coarse margin below its k-th entry too.  ids and distance bits equal the
    oracle's for one query tile (single CTAs) and for CTA pairs; k > 64 keeps the exact fp32 scan."""
    n = 30_000
    X = synth100k.table(n)
    Q, _ = synth100k.queries(200, n)
    ids = _ids(n)
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    with Index(dtype, capacity=n) as ix:
        ix.upsert(ids, X)
        for k, nq, path in ((16, 9, 2), (17, 7, 2), (32, 200, 2), (33, 130, 2), (64, 200, 2), (64, 3, 2), (65, 3, 1), (128, 5, 1)):
            g_ids, g_d, g_c = ix.search(Q[:nq], k)
            assert ix.stats()["last_path"] == path, (k, nq)
            for i in range(0, nq, 1 if nq < 10 else 9):
                w_ids, w_d = O.topk_exact(rows, ids, Q[i], k)
                assert g_c[i] == k and np.array_equal(g_ids[i], w_ids), (k, nq, i)
                assert np.array_equal(g_d[i].view(np.uint64), w_d.view(np.uint64)), (k, nq, i)
        assert ix.stats()["fallback_exhaustive"] == 0
    with Index(dtype) as ix:                                        # fewer live rows than k
        ix.upsert(ids[:4200], X[:4200])
        ix.delete(ids[100:4190])
        g_ids, g_d, g_c = ix.search(Q[:4], 64)
        keep = np.r_[0:100, 4190:4200]
        for i in range(4):
            w_ids, w_d = O.topk_exact(rows[keep], ids[keep], Q[i], 64)
            assert g_c[i] == 64 and np.array_equal(g_ids[i], w_ids) and np.array_equal(g_d[i].view(np.uint64), w_d.view(np.uint64))
        g_ids, g_d, g_c = ix.search(Q[:4], 128)
        for i in range(4):
            w_ids, w_d = O.topk_exact(rows[keep], ids[keep], Q[i], 128)
            assert g_c[i] == 110 and np.array_equal(g_ids[i, :110], w_ids) and np.isnan(g_d[i, 110:]).all()


def test_large_k_for_a_wider_reranker_feed(Index, small_table):
    """k up to 128: batches stay exact (k <= 64: tcgen05 scan with wider candidate lists, a tie wider than
    a list falls back to the fp32 scan; beyond: fp32 scan per query), ties and short tables too; the sharded exchange
    chunks the batch so that a slot still fits."""
    X, Q, _ = small_table
    n = 6000
    Xd = X[:n].copy()
    Xd[100:260] = Xd[5]                                        # 160 identical rows: wider than k
    ids = _ids(n)
    with Index("fp32") as ix:
        ix.upsert(ids, Xd)
        _check_exact(ix, Xd, ids, np.concatenate([Q[:5], Xd[5:6]]), k=100)
        assert ix.stats()["last_path"] == 1                    # k > 64: one fp32 scan per query
        _check_exact(ix, Xd, ids, np.concatenate([Q[:5], Xd[5:6]]), k=60)
        assert ix.stats()["last_path"] == 2                    # k <= 64: tcgen05 scan with 160-key candidate lists
        st = ix.stats()
        assert st["fallback_gemv"] + st["fallback_exhaustive"] >= 1     # the 161-row tie floods a list: exact re-answer
        # neighbouring flagged queries are re-answered together (one batched fp32 pass for the run)
        _check_exact(ix, Xd, ids, np.concatenate([Xd[5:6], Xd[5:6] * np.float32(2), Xd[100:101], Q[:2]]), k=60)
        assert ix.stats()["fallback_gemv"] >= st["fallback_gemv"] + 3
        ix.shard_connect([ix.shard_export(1, 0)])
        a = ix.search(Q[:9], 128)
        b = ix.search_sharded(Q[:9], 128)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))
        # two disjoint 64-entry lists (even / odd ranks of the top-128) merge back to the top-64
        m = ix.merge_topk(np.stack([a[0][:, 0::2], a[0][:, 1::2]]), np.stack([a[1][:, 0::2], a[1][:, 1::2]]),
                          np.stack([np.full(9, 64, np.int32)] * 2), 64)
        assert np.array_equal(m[0], a[0][:, :64]) and np.array_equal(m[1].view(np.uint64), a[1][:, :64].view(np.uint64))
    with Index("bf16") as ix:
        ix.upsert(ids[:50], Xd[:50])
        g = ix.search(Q[:2], 128)
        assert (g[2] == 50).all()


def test_two_searches_in_flight_equal_the_synchronous_answers(Index, small_table):
    """orx_search_submit / orx_search_wait: two complete scratch sets, so query i+1 is launched while query i is still in
    flight.  Same answers as orx_search (bit for bit), also when a query needs the exact fallback, when writes land
    between submit and wait of other tickets, and the bookkeeping errors are reported."""
    import torch
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    ids = _ids(X.shape[0])
    with Index("fp32") as ix:
        ix.upsert(ids, X)
        dup = np.tile(X[7], (300, 1))                               # a flood of exact ties: query 7's row needs the exact path
        ix.upsert(_ids(300, 500_000), dup)
        allX, all_ids = np.concatenate([X, dup]), np.concatenate([ids, _ids(300, 500_000)])
        Qd = torch.from_numpy(np.concatenate([Q[:6], X[7:8]])).cuda()
        want = [ix.search(Qd[i:i + 1].cpu().numpy(), K) for i in range(7)]
        outs = [(torch.empty((1, K, 2), dtype=torch.int64, device="cuda"), torch.empty((1, K), dtype=torch.float64, device="cuda"),
                 torch.empty((1,), dtype=torch.int32, device="cuda")) for _ in range(2)]
        prev = None
        got = {}
        for i in range(7):
            t = ix.search_submit(Qd[i:i + 1], K, outs[i & 1])
            if prev is not None:
                ix.search_wait(prev[0])
                got[prev[1]] = tuple(o.cpu().numpy() for o in outs[prev[1] & 1])
            prev = (t, i)
        ix.search_wait(prev[0])
        got[prev[1]] = tuple(o.cpu().numpy() for o in outs[prev[1] & 1])
        for i in range(7):
            assert np.array_equal(got[i][0].view(np.uint64), want[i][0]), i
            assert np.array_equal(got[i][1].view(np.uint64), want[i][1].view(np.uint64)), i
        w_ids, _ = O.topk_exact(allX, all_ids, X[7], K)
        assert np.array_equal(got[6][0][0].view(np.uint64), w_ids) and ix.stats()["fallback_exhaustive"] > 0
        # batches (tcgen05 path) in flight, a write between the submits: each search sees the table of its submit time
        Qb = torch.from_numpy(Q[:24]).cuda()
        ob = [(torch.empty((24, K, 2), dtype=torch.int64, device="cuda"), torch.empty((24, K), dtype=torch.float64, device="cuda"),
               torch.empty((24,), dtype=torch.int32, device="cuda")) for _ in range(2)]
        before = ix.search(Q[:24], K)
        t0 = ix.search_submit(Qb, K, ob[0])
        victims = before[0][:, 0].copy()
        assert ix.delete(victims) == len({tuple(v) for v in victims.tolist()})
        t1 = ix.search_submit(Qb, K, ob[1])
        ix.search_wait(t0)
        ix.search_wait(t1)
        assert np.array_equal(ob[0][0].cpu().numpy().view(np.uint64), before[0])
        after = ix.search(Q[:24], K)
        assert np.array_equal(ob[1][0].cpu().numpy().view(np.uint64), after[0])
        assert not np.isin(after[0][:, :, 1], victims[:, 1]).any()
        # bookkeeping
        a = ix.search_submit(Qd[:1], K, outs[0])
        b = ix.search_submit(Qd[1:2], K, outs[1])
        with pytest.raises(orx.OrxValueError, match="two searches are in flight"):
            ix.search_submit(Qd[2:3], K, outs[0])
        with pytest.raises(orx.OrxValueError, match="in flight"):
            ix.search_filtered(Q[:1], K, ids[:10])
        ix.search_wait(a)
        with pytest.raises(orx.OrxValueError, match="not in flight"):
            ix.search_wait(a)
        ix.search_wait(b)
        assert ix.search(Q[:1], K)[2][0] == K
