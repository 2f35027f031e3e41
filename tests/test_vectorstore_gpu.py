"""The drop-in `GpuVectorStore` driven the way the reference drives `AsyncPGVectorStore`:
create -> as_retriever(k=TOP_K) -> aadd_documents / adelete in rag.py's order (delete old chunk
ids of the batch's source docs, then add the re-chunked ones; reference app/rag.py:216-235)."""
import asyncio

import numpy as np
import pytest

from oracle import cosine_topk as O

pytestmark = pytest.mark.gpu


class FakeBgeM3:
    """Stand-in for the remote bge-m3 service (reference app/llm_services.py:218-222): the text
    IS the row index of the synthetic table / query set."""

    def __init__(self, syn, n_rows):
        self.syn, self.n_rows = syn, n_rows

    def embed_documents(self, texts):
        return self.syn.rows(np.array([int(t.split(":")[1]) for t in texts], np.uint64)).tolist()

    def embed_query(self, text):
        return self.syn.queries(1, self.n_rows, start=int(text.split(":")[1]))[0][0].tolist()

    async def aembed_documents(self, texts):
        return self.embed_documents(texts)

    async def aembed_query(self, text):
        return self.embed_query(text)


def test_reference_call_sequence(synth100k):
    import uuid
    import outline_rag_b200 as orx
    n_docs, per_doc = 60, 20
    n = n_docs * per_doc
    emb = FakeBgeM3(synth100k, n)

    async def run():
        store = await orx.GpuVectorStore.create(
            engine=None, embedding_service=emb, table_name="langchain_pg_embedding",
            metadata_columns=["source_id", "title", "outline_updated_at_str", "url"])
        retriever = store.as_retriever(search_kwargs={"k": orx.TOP_K})
        docs = [orx.Document(page_content=f"row:{i}", metadata={"source_id": f"doc{i // per_doc}", "title": "t",
                                                                "outline_updated_at_str": "2026", "url": "/d"},
                             id=str(uuid.UUID(int=i))) for i in range(n)]
        out_ids = await store.aadd_documents(docs)
        assert out_ids == [d.id for d in docs] and len(store.index) == n
        hits = await retriever.ainvoke("q:0")
        assert len(hits) == orx.TOP_K and all(isinstance(h, orx.Document) for h in hits)
        scored = await store.asimilarity_search_with_score_by_vector(emb.embed_query("q:0"), k=orx.TOP_K)

        # webhook refresh of 3 docs: look up old chunk ids by source_id, delete, add re-chunked rows
        stale = store.doc_store.ids_for_source(["doc0", "doc1", "doc2"])
        assert len(stale) == 3 * per_doc
        assert await store.adelete(ids=stale) is True
        assert len(store.index) == n - 3 * per_doc
        fresh = [orx.Document(page_content=f"row:{n + i}", metadata={"source_id": f"doc{i // per_doc}"})
                 for i in range(3 * per_doc)]
        new_ids = await store.aadd_documents(fresh)                   # no ids -> uuid4()
        assert len(set(new_ids)) == len(fresh) and len(store.index) == n
        hits2 = await retriever.ainvoke("q:0")
        assert not ({h.id for h in hits2} & set(stale))
        assert await store.adelete(ids=[]) is False
        store.index.close()
        return hits, scored

    hits, scored = asyncio.run(run())
    X = synth100k.table(n)
    q = np.asarray(emb.embed_query("q:0"), np.float32)
    w_ids, w_d = O.topk_exact(X, O.ids_arange(0, n), q, 12)
    import uuid
    assert [h.id for h in hits] == [str(uuid.UUID(int=v)) for v in O.ids_to_ints(w_ids)]
    assert [h.page_content for h in hits] == [f"row:{v}" for v in O.ids_to_ints(w_ids)]
    assert hits[0].metadata["source_id"] == f"doc{O.ids_to_ints(w_ids)[0] // 20}"
    assert [s for _, s in scored] == w_d.tolist()                      # cosine DISTANCE, ascending


def test_errors_propagate_like_the_reference(synth100k):
    import outline_rag_b200 as orx

    class BadDim(FakeBgeM3):
        def embed_documents(self, texts):
            return [[0.0] * 768 for _ in texts]

    store = orx.GpuVectorStore.create_sync(BadDim(synth100k, 10))
    with pytest.raises(ValueError, match="dimensions"):
        store.add_documents([orx.Document(page_content="row:1")])
    assert len(store.index) == 0
    store.index.close()


def test_metadata_filter_restricts_the_candidates(synth100k):
    """`filter=` (upstream VectorStore argument): predicate -> ids in the doc store, ordering on the GPU."""
    import uuid
    import outline_rag_b200 as orx
    n, per_doc = 2000, 20
    emb = FakeBgeM3(synth100k, n)
    store = orx.GpuVectorStore.create_sync(emb)
    docs = [orx.Document(page_content=f"row:{i}", metadata={"source_id": f"doc{i // per_doc}", "title": f"t{i % 3}"},
                         id=str(uuid.UUID(int=i))) for i in range(n)]
    store.add_documents(docs)
    q = emb.embed_query("q:3")
    X = synth100k.table(n)
    ids = O.ids_arange(0, n)
    for flt, keep in [({"source_id": "doc7"}, [i for i in range(n) if i // per_doc == 7]),
                      ({"source_id": {"$in": ["doc1", "doc2", "doc50"]}}, [i for i in range(n) if i // per_doc in (1, 2, 50)]),
                      ({"$and": [{"source_id": {"$in": ["doc1", "doc2"]}}, {"title": "t0"}]},
                       [i for i in range(n) if i // per_doc in (1, 2) and i % 3 == 0]),
                      ({"source_id": "no-such-doc"}, [])]:
        got = store.similarity_search_with_score_by_vector(q, k=12, filter=flt)
        keep = np.asarray(keep, np.int64)
        w_ids, w_d = O.topk_exact(X[keep], ids[keep], np.asarray(q, np.float32), 12, exhaustive=True)
        assert [d.id for d, _ in got] == [str(uuid.UUID(int=v)) for v in O.ids_to_ints(w_ids)]
        assert [s for _, s in got] == w_d.tolist()
    store.index.close()


def test_restart_rebuilds_the_device_table_from_the_doc_store(synth100k):
    """Postgres stays the source of truth: a new process streams `COPY (SELECT langchain_id, embedding ...)` from
    it (here: the in-memory stand-in's `copy_binary`) and answers exactly like the process that was stopped."""
    import uuid
    import outline_rag_b200 as orx
    n = 500
    emb = FakeBgeM3(synth100k, n)

    async def run():
        first = await orx.GpuVectorStore.create(None, emb)
        docs = [orx.Document(page_content=f"row:{i}", metadata={"source_id": f"doc{i // 20}"}, id=str(uuid.UUID(int=i + 1)))
                for i in range(n)]
        await first.aadd_documents(docs)
        assert await first.adelete(ids=[d.id for d in docs[40:60]]) is True
        before = [await first.asimilarity_search_with_score(f"q:{j}", k=orx.TOP_K) for j in range(4)]
        durable = first.doc_store                                  # survives the "restart"
        first.index.close()

        second = await orx.GpuVectorStore.create(None, emb, doc_store=durable)

        async def copy_stream():
            for chunk in durable.copy_binary(rows_per_chunk=64):
                yield chunk

        assert await second.aload_pgcopy(copy_stream(), feed_bytes=1 << 20) == (n - 20, 0)
        after = [await second.asimilarity_search_with_score(f"q:{j}", k=orx.TOP_K) for j in range(4)]
        for b, a in zip(before, after):
            assert [(d.id, d.page_content, s) for d, s in b] == [(d.id, d.page_content, s) for d, s in a]
        second.index.close()

    asyncio.run(run())
