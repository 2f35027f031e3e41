"""The reference's seam is TYPED (reference app/rag.py:28-31, :85-99): `vector_store.as_retriever(search_kwargs={"k": 12})`
must be a langchain `BaseRetriever` / Runnable, because it is handed to the pydantic-validated
`ContextualCompressionRetriever(base_compressor=..., base_retriever=base_retriever)`.  langchain is not installed in this
image, so the check runs against minimal stand-ins of langchain-core's classes (tests/stubs/, shaped after 0.3.x) in a
fresh interpreter: with `langchain_core` importable `GpuVectorStore` must BE a `VectorStore` and its retriever a real
`VectorStoreRetriever`; without it the duck-typed fallback is used.  No GPU: the index is a fake with the `Index`
methods the store calls."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import asyncio, sys, uuid
    import numpy as np
    sys.path.insert(0, {stubs!r})
    sys.path.insert(0, {root!r})
    import langchain_core
    from langchain_core.documents import Document
    from langchain_core.retrievers import BaseRetriever
    from langchain_core.runnables import Runnable
    from langchain_core.vectorstores import VectorStore, VectorStoreRetriever
    from langchain.retrievers.contextual_compression import ContextualCompressionRetriever
    import pydantic
    import outline_rag_b200 as orx
    from outline_rag_b200 import vectorstore as vs
    assert vs.HAVE_LANGCHAIN and vs.Document is Document

    class FakeIndex:                                   # the Index surface GpuVectorStore touches
        def __init__(self): self.rows = {{}}
        def upsert(self, ids, vecs):
            for i, v in zip(ids, np.asarray(vecs, np.float32)): self.rows[str(uuid.UUID(str(i)))] = v
        def delete(self, ids):
            for i in ids: self.rows.pop(str(uuid.UUID(str(i))), None)
        def search(self, q, k):
            q = np.asarray(q, np.float32).reshape(-1, 1024)
            keys = sorted(self.rows)
            X = np.stack([self.rows[i] for i in keys]) if keys else np.zeros((0, 1024), np.float32)
            ids = np.zeros((q.shape[0], k, 2), np.uint64); dist = np.full((q.shape[0], k), np.nan); cnt = np.zeros(q.shape[0], np.int32)
            for j in range(q.shape[0]):
                d = 1.0 - (X @ q[j]) / (np.linalg.norm(X, axis=1) * np.linalg.norm(q[j]))
                order = np.argsort(d, kind="stable")[:k]
                for r, o in enumerate(order):
                    v = uuid.UUID(keys[o]).int
                    ids[j, r] = (v >> 64, v & (2**64 - 1)); dist[j, r] = d[o]
                cnt[j] = len(order)
            return ids, dist, cnt

    class Emb:
        def _e(self, t):
            rng = np.random.default_rng(abs(hash(t)) % 2**32); return rng.standard_normal(1024).astype(np.float32)
        def embed_documents(self, texts): return [self._e(t) for t in texts]
        def embed_query(self, t): return self._e(t)
        async def aembed_documents(self, texts): return self.embed_documents(texts)
        async def aembed_query(self, t): return self.embed_query(t)

    class Reranker:                                    # stands in for the remote bge-reranker compressor
        async def acompress_documents(self, docs, query): return docs[:3]
        def compress_documents(self, docs, query): return docs[:3]

    store = orx.GpuVectorStore(FakeIndex(), Emb())
    assert isinstance(store, VectorStore)                                     # rag.py:28 `Optional[AsyncPGVectorStore]`
    docs = [Document(page_content=f"chunk {{i}}", metadata={{"source_id": f"d{{i // 4}}", "title": "t", "url": "u",
                     "outline_updated_at_str": "x"}}, id=str(uuid.UUID(int=i + 1))) for i in range(20)]

    async def main():
        ids = await store.aadd_documents(docs)                                # rag.py:235
        assert ids == [d.id for d in docs]
        base = store.as_retriever(search_kwargs={{"k": 12}})                   # rag.py:85-87
        assert isinstance(base, VectorStoreRetriever) and isinstance(base, BaseRetriever) and isinstance(base, Runnable)
        assert base.search_kwargs == {{"k": 12}} and base.vectorstore is store
        comp = ContextualCompressionRetriever(base_compressor=Reranker(), base_retriever=base)      # rag.py:96-99
        got = await base.ainvoke("chunk 7")
        assert len(got) == 12 and all(isinstance(d, Document) for d in got) and got[0].page_content == "chunk 7"
        assert got[0].metadata["source_id"] == "d1" and got[0].id == docs[7].id
        top = await comp.ainvoke("chunk 7")                                   # api.py:122
        assert [d.page_content for d in top] == [d.page_content for d in got[:3]]
        assert await store.adelete(ids=[docs[7].id]) is True                  # rag.py:231
        again = await base.ainvoke("chunk 7")
        assert all(d.id != docs[7].id for d in again) and len(again) == 12
        assert base.invoke("chunk 3")[0].page_content == "chunk 3"            # the sync path of the same retriever
        # the duck-typed stand-in is NOT accepted by the reference's constructor -- the reason for all of the above
        try:
            ContextualCompressionRetriever(base_compressor=Reranker(), base_retriever=vs.GpuRetriever(store, {{"k": 12}}))
        except pydantic.ValidationError:
            pass
        else:
            raise AssertionError("a plain class passed the RetrieverLike check")
        # langchain's generic entry points land on the same code
        more = store.add_texts(["late chunk"], [{{"source_id": "z"}}])
        assert len(more) == 1 and (await store.asimilarity_search("late chunk", k=1))[0].page_content == "late chunk"
    asyncio.run(main())
    print("SEAM-OK")
''')


def test_store_and_retriever_are_langchain_types_when_langchain_is_present():
    code = SCRIPT.format(stubs=os.path.join(ROOT, "tests", "stubs"), root=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and "SEAM-OK" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]


def test_duck_typed_fallback_without_langchain():
    from outline_rag_b200 import vectorstore as vs
    if vs.HAVE_LANGCHAIN:          # a host with langchain installed: the typed path above is the one that matters
        return
    r = vs.GpuVectorStore.__mro__
    assert r[1] is object
    store = vs.GpuVectorStore(index=None, embedding_service=object())
    retr = store.as_retriever(search_kwargs={"k": 12})
    assert isinstance(retr, vs.GpuRetriever) and retr.search_kwargs == {"k": 12}
