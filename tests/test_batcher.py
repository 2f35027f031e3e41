"""Micro-batching front end (outline_rag_b200/batcher.py).  Host logic on CPU with a fake index;
on the GPU, concurrent requests through `GpuVectorStore` must equal one-at-a-time results while
issuing far fewer scans."""
import asyncio

import numpy as np
import pytest

from oracle import cosine_topk as O


class FakeIndex:
    """Index stand-in (oracle-backed) that records the batches it was handed."""

    def __init__(self, X, ids):
        self.X, self.ids, self.calls = X, ids, []

    def search(self, Q, k):
        self.calls.append((Q.shape[0], k))
        nq = Q.shape[0]
        ids = np.zeros((nq, k, 2), np.uint64)
        d = np.full((nq, k), np.nan)
        c = np.zeros(nq, np.int32)
        for i in range(nq):
            a, b = O.topk_exact(self.X, self.ids, Q[i], k)
            ids[i, :len(b)], d[i, :len(b)], c[i] = a, b, len(b)
        return ids, d, c


def test_coalesces_concurrent_requests_and_keeps_per_request_semantics(small_table):
    from outline_rag_b200.batcher import QueryBatcher
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    X = X[:600]
    ids = O.ids_arange(0, 600)
    fake = FakeIndex(X, ids)

    async def run():
        b = QueryBatcher(fake, max_batch=8, max_wait_ms=20)
        ks = [12, 3, 12, 1, 5, 12, 7, 12, 2, 12, 12]            # 11 requests: one full batch of 8 + 3
        bad = Q[0].copy()
        bad[5] = np.nan
        tasks = [asyncio.create_task(b.search(Q[i], k)) for i, k in enumerate(ks)]
        tasks.append(asyncio.create_task(b.search(bad, 12)))     # rejected alone
        tasks.append(asyncio.create_task(b.search(Q[0][:100], 12)))
        res = await asyncio.gather(*tasks, return_exceptions=True)
        await b.drain()
        return b, res

    b, res = asyncio.run(run())
    ks = [12, 3, 12, 1, 5, 12, 7, 12, 2, 12, 12]
    for i, k in enumerate(ks):
        w_ids, w_d = O.topk_exact(X, ids, Q[i], k)
        assert np.array_equal(res[i][0], w_ids) and np.array_equal(res[i][1], w_d)
    assert isinstance(res[11], orx.OrxValueError) and "NaN or infinite" in str(res[11])
    assert isinstance(res[12], orx.OrxValueError) and "dimensions" in str(res[12])
    assert fake.calls == [(8, 12), (3, 12)] and b.batches == 2 and b.requests == 11


def test_a_window_is_split_by_filter_handle_and_k_class(small_table):
    """One engine call serves the requests that share a prepared filter (or none) and a k class (<= 16, <= 64, beyond:
    the candidate-list widths of the batched scan), so a k = 100 request or a filtered one never changes the path the
    k = 12 requests of the same window take.  Every caller still gets exactly its own top-k."""
    from outline_rag_b200.batcher import QueryBatcher
    X, Q, _ = small_table
    X = X[:400]
    ids = O.ids_arange(0, 400)

    class Flt:                                                  # stands in for engine.Filter (a device-resident handle)
        def __init__(self, rows):
            self.rows = np.asarray(rows)

    class Fake(FakeIndex):
        def search_filtered(self, Q, k, flt):
            self.calls.append((Q.shape[0], k, "filtered", len(flt.rows)))
            sub = FakeIndex(self.X[flt.rows], self.ids[flt.rows])
            return sub.search(Q, k)

    fake = Fake(X, ids)
    f_a, f_b = Flt(np.arange(0, 400, 2)), Flt(np.arange(100, 300))
    reqs = [(12, None), (100, None), (5, None), (40, None), (12, f_a), (3, f_a), (12, f_b), (64, None), (16, None), (70, f_a)]

    async def run():
        b = QueryBatcher(fake, max_batch=64, max_wait_ms=20)
        res = await asyncio.gather(*[b.search(Q[i], k, flt) for i, (k, flt) in enumerate(reqs)])
        await b.drain()
        return b, res

    b, res = asyncio.run(run())
    for i, (k, flt) in enumerate(reqs):
        rows = np.arange(400) if flt is None else flt.rows
        w_ids, w_d = O.topk_exact(X[rows], ids[rows], Q[i], k)
        assert np.array_equal(res[i][0], w_ids) and np.array_equal(res[i][1], w_d), i
    assert sorted(fake.calls, key=str) == sorted([(3, 16), (2, 64), (1, 100), (2, 12, "filtered", 200), (1, 12, "filtered", 200),
                                                 (1, 70, "filtered", 200)], key=str)
    assert b.batches == 6 and b.requests == len(reqs)


def test_window_flushes_a_lone_request_and_engine_errors_reach_every_waiter(small_table):
    from outline_rag_b200.batcher import QueryBatcher
    X, Q, _ = small_table

    class Broken:
        def search(self, Q, k):
            raise RuntimeError("device lost")

    async def run():
        ok = QueryBatcher(FakeIndex(X[:100], O.ids_arange(0, 100)), max_batch=64, max_wait_ms=1)
        one = await asyncio.wait_for(ok.search(Q[1], 4), timeout=5)
        bad = QueryBatcher(Broken(), max_batch=64, max_wait_ms=1)
        errs = await asyncio.gather(bad.search(Q[0], 4), bad.search(Q[1], 4), return_exceptions=True)
        return one, errs

    one, errs = asyncio.run(run())
    assert one[0].shape == (4, 2)
    assert all(isinstance(e, RuntimeError) and "device lost" in str(e) for e in errs)


@pytest.mark.gpu
def test_vectorstore_with_batching_on_gpu(synth100k):
    import uuid
    import outline_rag_b200 as orx
    from tests.test_vectorstore_gpu import FakeBgeM3
    n = 8192
    emb = FakeBgeM3(synth100k, n)

    async def run():
        store = await orx.GpuVectorStore.create(embedding_service=emb, batch_window_ms=5.0, max_batch=64)
        docs = [orx.Document(page_content=f"row:{i}", metadata={"source_id": f"d{i // 16}"}, id=str(uuid.UUID(int=i)))
                for i in range(n)]
        await store.aadd_documents(docs)
        retriever = store.as_retriever(search_kwargs={"k": orx.TOP_K})
        hits = await asyncio.gather(*[retriever.ainvoke(f"q:{i}") for i in range(40)])
        await store.batcher.drain()
        stats = store.index.stats()
        batches = store.batcher.batches
        store.index.close()
        return hits, stats, batches

    hits, stats, batches = asyncio.run(run())
    X = synth100k.table(n)
    ids = O.ids_arange(0, n)
    for i in range(40):
        q = np.asarray(emb.embed_query(f"q:{i}"), np.float32)
        w_ids, _ = O.topk_exact(X, ids, q, 12)
        assert [h.id for h in hits[i]] == [str(uuid.UUID(int=v)) for v in O.ids_to_ints(w_ids)]
    assert batches <= 4 and stats["last_path"] == 2           # 40 requests -> a handful of tcgen05 scans


@pytest.mark.gpu
def test_requests_sharing_a_prepared_filter_share_one_filtered_pass(synth100k):
    """A source-scoped assistant asks every question under the same prepared filter: the concurrent requests of a window
    become ONE `search_filtered` call (one tcgen05 pass with the predicate folded into the row scale), unfiltered
    requests of the same window keep their own call, and every answer equals the oracle's on its own row set."""
    import uuid
    import outline_rag_b200 as orx
    from tests.test_vectorstore_gpu import FakeBgeM3
    n = 8192
    emb = FakeBgeM3(synth100k, n)

    async def run():
        store = await orx.GpuVectorStore.create(embedding_service=emb, batch_window_ms=20.0, max_batch=64)
        docs = [orx.Document(page_content=f"row:{i}", metadata={"source_id": f"d{i % 2}"}, id=str(uuid.UUID(int=i)))
                for i in range(n)]
        await store.aadd_documents(docs)
        flt = store.prepare_filter({"source_id": "d1"})              # the odd rows: 4096 eligible -> bitmap regime
        filtered = [store.asimilarity_search(f"q:{i}", k=orx.TOP_K, filter=flt) for i in range(12)]
        plain = [store.asimilarity_search(f"q:{i}", k=orx.TOP_K) for i in range(12, 18)]
        hits = await asyncio.gather(*filtered, *plain)
        await store.batcher.drain()
        batches, requests = store.batcher.batches, store.batcher.requests
        flt.close()
        store.index.close()
        return hits, batches, requests

    hits, batches, requests = asyncio.run(run())
    X = synth100k.table(n)
    ids = O.ids_arange(0, n)
    odd = np.arange(1, n, 2)
    for i in range(18):
        q = np.asarray(emb.embed_query(f"q:{i}"), np.float32)
        rows = odd if i < 12 else np.arange(n)
        w_ids, _ = O.topk_exact(X[rows], ids[rows], q, 12)
        assert [h.id for h in hits[i]] == [str(uuid.UUID(int=v)) for v in O.ids_to_ints(w_ids)], i
    assert requests == 18 and batches == 2                         # one filtered call + one plain call
