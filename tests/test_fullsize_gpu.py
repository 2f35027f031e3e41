"""Parity at 1M rows x 1024 (BASELINE.json configs[1]) against a TRUE oracle pass over the whole table.

The table is built in HBM by the device generator; the SAME rows are generated on the host by the bit-identical C
generator (oracle/synth_host.c) and scanned completely by the oracle (oracle/cosine_topk.py StreamingTopK: fp32 BLAS
shortlist per chunk with a rigorous margin, canonical binary64 rescoring).  The engine's ids AND float64 distance bits
must equal the oracle's for every query, on the GEMV scan (single queries) and on both tcgen05 scans (batches), for
fp32 tables and for bf16 tables (oracle fed with the rows as a bf16 table stores them).  Plus the size-independent
properties: sortedness, self-query, delete-shifts-the-ranking / reinsert-restores-it, bf16 recall@12.
"""
import numpy as np
import pytest

from oracle import cosine_topk as O
from tests._helpers import stored_bf16_rows

pytestmark = pytest.mark.gpu
N, K, NQ = 1_000_000, 12, 48


@pytest.fixture(scope="module")
def big():
    import outline_rag_b200 as orx
    from oracle.synth_host import FastSynth
    from orx_testkit.device import synth_rows_device
    from orx_testkit.synth import SEED_TABLE, default_centres
    syn = FastSynth(default_centres(N))
    tables = {}
    for dtype in ("fp32", "bf16"):
        ix = orx.Index(dtype, N + 1024, 0)
        chunk = 262_144
        for s in range(0, N, chunk):
            m = min(chunk, N - s)
            rows = synth_rows_device(0, SEED_TABLE, syn.n_centres, s, m)
            ids = np.zeros((m, 2), np.uint64)
            ids[:, 1] = np.arange(s, s + m, dtype=np.uint64)
            ix.upsert(ids, rows)
        tables[dtype] = ix
    Q, anchors = syn.queries(NQ, N)
    # ---- the oracle's full scan of the same table, generated independently on the host
    truth = {"fp32": O.StreamingTopK(Q, K), "bf16": O.StreamingTopK(Q, K)}
    for s in range(0, N, 65_536):
        m = min(65_536, N - s)
        X = syn.table(m, start=s)
        ids = O.ids_arange(s, s + m)
        truth["fp32"].feed(X, ids)
        truth["bf16"].feed(stored_bf16_rows(X), ids)
    want = {d: t.result() for d, t in truth.items()}
    assert truth["fp32"].rows_seen == N
    yield syn, tables, Q, anchors, want
    for ix in tables.values():
        ix.close()


def _same(got_ids, got_d, want_pair):
    return np.array_equal(got_ids, want_pair[0]) and np.array_equal(got_d.view(np.uint64), want_pair[1].view(np.uint64))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_every_scan_path_equals_the_full_oracle_pass(big, dtype):
    syn, tables, Q, anchors, want = big
    ix = tables[dtype]
    # tcgen05 scan on CTA pairs (more than 128 queries is not needed for that: > TILE_M selects pairs) -- 48 queries
    # run on one CTA per query tile; 48 + 96 repeated = 144 queries run on pairs
    batch = ix.search(Q, K)
    assert ix.stats()["last_path"] == 2
    for i in range(NQ):
        assert _same(batch[0][i], batch[1][i], want[dtype][i]), (dtype, "tcgen05 1-CTA", i)
    Q3 = np.concatenate([Q, Q, Q])
    pairs = ix.search(Q3, K)
    assert ix.stats()["last_path"] == 2
    for i in range(3 * NQ):
        assert _same(pairs[0][i], pairs[1][i], want[dtype][i % NQ]), (dtype, "tcgen05 CTA pairs", i)
    for i in range(0, NQ, 3):
        one = ix.search(Q[i], K)                                  # GEMV scan
        assert ix.stats()["last_path"] == 1
        assert _same(one[0][0], one[1][0], want[dtype][i]), (dtype, "gemv", i)
    assert (batch[2] == K).all()
    assert (np.diff(batch[1], axis=1) >= 0).all()
    assert (batch[0][:, 0, 1] == anchors.astype(np.uint64)).all()       # the anchor row is the best hit
    st = ix.stats()
    assert st["fallback_exhaustive"] == 0                              # the fast paths proved these answers themselves


def test_larger_k_equals_the_full_oracle_pass(big):
    """k = 100 (a wider reranker feed) on the fp32 table, against its own full oracle scan."""
    syn, tables, Q, _, _ = big
    k = 100
    st = O.StreamingTopK(Q[:3], k)
    for s in range(0, N, 131_072):
        m = min(131_072, N - s)
        st.feed(syn.table(m, start=s), O.ids_arange(s, s + m))
    got = tables["fp32"].search(Q[:3], k)
    for i, w in enumerate(st.result()):
        assert _same(got[0][i], got[1][i], w), i


def test_self_query_and_bf16_recall(big):
    syn, tables, Q, _, want = big
    rows = np.array([5, 123_456, 999_999], np.uint64)
    ids, dist, _ = tables["fp32"].search(syn.rows(rows), 1)
    assert (ids[:, 0, 1] == rows).all() and (np.abs(dist[:, 0]) < 1e-15).all()
    recall = np.mean([len(set(f[0][:, 1].tolist()) & set(b[0][:, 1].tolist())) / K
                      for f, b in zip(want["fp32"], want["bf16"])])
    assert recall >= 0.99, recall                                       # oracle vs oracle: a property of bf16 storage
    f = tables["fp32"].search(Q, K)[0][:, :, 1]
    b = tables["bf16"].search(Q, K)[0][:, :, 1]
    assert np.mean([len(set(x) & set(y)) / K for x, y in zip(f.tolist(), b.tolist())]) == recall


def test_delete_shifts_the_ranking_and_reinsert_restores_it(big):
    syn, tables, Q, _, _ = big
    ix = tables["fp32"]
    before = ix.search(Q[:4], K + 1)
    victims = before[0][:, 0].copy()                              # every query's best hit
    assert ix.delete(victims) == 4 and len(ix) == N - 4
    after = ix.search(Q[:4], K)
    assert np.array_equal(after[0], before[0][:, 1:])             # ranks 2..13 move up unchanged
    assert np.array_equal(after[1].view(np.uint64), before[1][:, 1:].view(np.uint64))
    ix.upsert(victims, syn.rows(victims[:, 1]))
    again = ix.search(Q[:4], K + 1)
    assert np.array_equal(again[0], before[0]) and np.array_equal(again[1].view(np.uint64), before[1].view(np.uint64))
