"""The multi-GPU index inside ONE process (`orx_create_multi`, `Index(devices=[...])`, `GpuVectorStore.create(devices=...)`).

Runs the REAL row-sharded chain -- per-shard scan, finalize pushing each shard's candidates into the root's gather
buffer, merge_wait over `world > 1` slots, completion word polled by the host -- in the driver's 1-GPU `-m gpu` pass by
listing one device several times (`devices=[0, 0, 0]`: three shards on one GPU share one stream, so the merge is ordered
behind their kernels and never waits for a kernel that cannot run).  With >= 2 GPUs the same tests run across devices
over NVLink peer memory.  Everything is compared with the oracle bit for bit, like the single-GPU tests.
"""
import numpy as np
import pytest

from oracle import cosine_topk as O
from tests._helpers import stored_bf16_rows

pytestmark = pytest.mark.gpu
K = 12


def _device_sets():
    import torch
    sets = [[0, 0, 0]]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append(list(range(min(n, 8))))
        sets.append([0, 1, 1, 0])
    return sets


def _exact(ix, rows, ids, Q, k=K):
    g_ids, g_d, g_c = ix.search(Q, k)
    for i in range(Q.shape[0]):
        w_ids, w_d = O.topk_exact(rows, ids, Q[i], k)
        m = len(w_d)
        assert g_c[i] == m, (i, g_c[i], m)
        assert np.array_equal(g_ids[i, :m], w_ids), f"query {i}: ids differ"
        assert np.array_equal(g_d[i, :m].view(np.uint64), w_d.view(np.uint64)), f"query {i}: distance bits differ"
    return g_ids, g_d, g_c


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_group_equals_the_oracle_on_every_scan_path(synth100k, dtype):
    import torch
    import outline_rag_b200 as orx
    n = 30_000
    X = synth100k.table(n)
    Q, _ = synth100k.queries(200, n)
    ids = O.ids_arange(0, n)
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    for devs in _device_sets():
        with orx.Index(dtype, capacity=n, devices=devs) as ix:
            assert ix.shard_count == len(devs)
            ix.upsert(ids[:20_000], X[:20_000])                                  # host rows
            ix.upsert(ids[20_000:], torch.from_numpy(X[20_000:]).cuda(devs[0]))  # device rows (gathered per shard)
            assert len(ix) == n
            _exact(ix, rows, ids, Q[:1])                          # every shard: GEMV scan
            _exact(ix, rows, ids, Q[:5])                          # every shard: tcgen05 scan
            g = _exact(ix, rows, ids, Q)                          # 200 queries: tcgen05 on CTA pairs
            # device-resident queries and results (on the first device)
            d = ix.search(torch.from_numpy(Q[:7]).cuda(devs[0]), K)
            assert np.array_equal(d[0].cpu().numpy().view(np.uint64), g[0][:7])
            assert np.array_equal(d[1].cpu().numpy().view(np.uint64), g[1][:7].view(np.uint64))
            st = ix.stats()
            assert st["searches"] == 4 and st["queries"] == 213 and st["kernel_launches"] > 0
            assert st["fallback_exhaustive"] == 0
            # k beyond the tensor path's candidate lists
            _exact(ix, rows, ids, Q[:2], k=50)


def test_group_write_path_filter_fetch_export(synth100k):
    import outline_rag_b200 as orx
    n = 12_000
    X = synth100k.table(n)
    Q, _ = synth100k.queries(6, n)
    ids = O.ids_arange(0, n)
    for devs in _device_sets():
        with orx.Index("fp32", devices=devs) as ix:
            ix.upsert(ids, X)
            # delete across shards; unknown ids are ignored
            gone = np.concatenate([ids[100:400], O.ids_arange(10**9, 10**9 + 5)])
            assert ix.delete(gone) == 300 and len(ix) == n - 300
            keep = np.ones(n, bool)
            keep[100:400] = False
            _exact(ix, X[keep], ids[keep], Q)
            assert ix.contains(5) and not ix.contains(150)
            # upsert replaces in place, last duplicate wins
            ix.upsert(np.concatenate([ids[5:6], ids[5:6]]), np.stack([X[9000], X[9001]]))
            X2 = X.copy()
            X2[5] = X[9001]
            _exact(ix, X2[keep], ids[keep], Q[:2])
            got, found = ix.fetch(np.concatenate([ids[5:7], ids[150:151]]))
            assert found.tolist() == [True, True, False] and np.array_equal(got[0], X[9001]) and np.array_equal(got[1], X[6])
            # a NaN anywhere rejects the whole batch, on every shard
            bad = X[:64].copy()
            bad[63, 1000] = np.nan
            with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
                ix.upsert(O.ids_arange(50_000, 50_064), bad)
            assert len(ix) == n - 300 and not ix.contains(50_000)
            with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
                q = Q[:2].copy()
                q[1, 0] = np.inf
                ix.search(q, K)
            # WHERE langchain_id IN (...): each shard answers for the ids it owns
            allow = ids[keep][::7]
            f_ids, f_d, f_c = ix.search_filtered(Q[:3], K, allow)
            sel = np.isin(np.arange(n), allow[:, 1].astype(np.int64))
            for i in range(3):
                w_ids, w_d = O.topk_exact(X2[sel], ids[sel], Q[i], K)
                assert np.array_equal(f_ids[i], w_ids) and np.array_equal(f_d[i].view(np.uint64), w_d.view(np.uint64))
            # export: every live row exactly once, verbatim; a fresh multi-GPU index rebuilt from it answers the same
            m = len(ix)
            e_ids = np.zeros((m, 2), np.uint64)
            e_rows = np.zeros((m, 4096), np.uint8)
            ix.export_rows(0, m, e_ids, e_rows)
            order = np.argsort(e_ids[:, 1])
            assert np.array_equal(e_ids[order], ids[keep])
            assert np.array_equal(e_rows.view(np.float32)[order], X2[keep])
            before = ix.search(Q, K)
        with orx.Index("fp32", devices=devs[::-1]) as ix2:
            ix2.upsert(e_ids, e_rows.view(np.float32))
            after = ix2.search(Q, K)
            assert np.array_equal(before[0], after[0]) and np.array_equal(before[1].view(np.uint64), after[1].view(np.uint64))


def test_group_unproven_queries_take_the_exact_path(synth100k):
    """More exact ties than a candidate list holds (and NaN rows, and fewer rows than k on some shards): the shards flag
    the query, the group re-answers it from every shard's exact search and merges on the root."""
    import outline_rag_b200 as orx
    n = 9_000
    X = synth100k.table(n).copy()
    X[11] = 0.0                                              # zero-norm row: distance NaN, sorts last
    ids = O.ids_arange(0, n)
    dup = np.tile(X[7], (300, 1))
    dup_ids = O.ids_arange(100_000, 100_300)
    for devs in _device_sets():
        with orx.Index("fp32", devices=devs) as ix:
            ix.upsert(ids, X)
            ix.upsert(dup_ids, dup)
            allX, all_ids = np.concatenate([X, dup]), np.concatenate([ids, dup_ids])
            g = _exact(ix, allX, all_ids, X[7:8], k=K)
            assert O.ids_to_ints(g[0][0])[0] == 7 and ix.stats()["fallback_exhaustive"] > 0
            g = _exact(ix, allX, all_ids, np.stack([X[7], X[100], X[7]]), k=K)      # a batch with flagged members
        with orx.Index("fp32", devices=devs) as tiny:          # fewer rows than k, some shards empty
            tiny.upsert(ids[:5], X[:5])
            g_ids, g_d, g_c = tiny.search(X[3:4], K)
            w_ids, w_d = O.topk_exact(X[:5], ids[:5], X[3], K)
            assert g_c[0] == 5 and np.array_equal(g_ids[0, :5], w_ids) and np.isnan(g_d[0, 5:]).all()
            assert tiny.search(X[:1], K)[2][0] == 5
        with orx.Index("fp32", devices=devs) as empty:
            assert empty.search(X[:2], K)[2].tolist() == [0, 0]


def test_vectorstore_over_a_multi_gpu_index(synth100k):
    """`GpuVectorStore.create(devices=[...])`: the reference's call sequence (rag.py:69-87, :231-235) on the group."""
    import asyncio
    import outline_rag_b200 as orx
    from tests.test_vectorstore_gpu import FakeBgeM3
    n = 3000
    emb = FakeBgeM3(synth100k, n)
    X = synth100k.table(n)

    async def main(devs):
        store = await orx.GpuVectorStore.create(None, emb, table_name="langchain_pg_embedding",
                                                metadata_columns=orx.vectorstore.DEFAULT_METADATA_COLUMNS, devices=devs)
        docs = [orx.Document(page_content=f"row:{i}", metadata={"source_id": f"doc{i // 20}", "title": "t"}) for i in range(n)]
        ids = await store.aadd_documents(docs)
        assert store.index.shard_count == len(devs) and len(store.index) == n
        retr = store.as_retriever(search_kwargs={"k": orx.TOP_K})
        got = await retr.ainvoke("query:3")
        _, anchors = synth100k.queries(4, n)
        assert len(got) == orx.TOP_K and got[0].page_content == f"row:{anchors[3]}"
        old = store.doc_store.ids_for_source([f"doc{anchors[3] // 20}"])
        assert await store.adelete(ids=old) is True and len(store.index) == n - 20
        again = await retr.ainvoke("query:3")
        assert all(d.metadata["source_id"] != f"doc{anchors[3] // 20}" for d in again)
        scored = await store.asimilarity_search_with_score_by_vector(X[77], k=3)
        assert scored[0][0].page_content == "row:77" and abs(scored[0][1]) < 1e-12
        store.index.close()

    for devs in _device_sets():
        asyncio.run(main(devs))


def test_group_cold_start_from_a_copy_stream(synth100k):
    """COPY BINARY cold start into a multi-GPU index: the loader stages and decodes on the first device, the upsert
    scatters the decoded rows to their shards; the table then answers like the oracle."""
    import outline_rag_b200 as orx
    from oracle import pgvector_wire as W
    n = 16384 + 3000                                            # two device batches
    X = synth100k.table(n)
    ids = O.ids_arange(0, n)
    stream = W.copy_binary_stream(ids, X)
    Q, _ = synth100k.queries(5, n)
    for devs in _device_sets():
        with orx.Index("fp32", devices=devs) as ix:
            with ix.pgcopy_loader() as ld:
                for o in range(0, len(stream), 3_000_017):
                    ld.feed(stream[o:o + 3_000_017])
            assert ld.result == (n, 0) and len(ix) == n
            _exact(ix, X, ids, Q[:1])
            _exact(ix, X, ids, Q)
