"""Host-side logic that needs no GPU: id conversions, the shard function, the doc-store filter
grammar, and `GpuVectorStore`'s call sequence against an oracle-backed fake index."""
import asyncio
import uuid

import numpy as np
import pytest

from oracle import cosine_topk as O
from tests.test_daemon import FakeOwner


def test_id_conversions_roundtrip_and_uuid_order():
    from outline_rag_b200.engine import ids_to_array, ids_to_ints, ids_to_uuid_strs
    vals = [0, 1, 2**64 - 1, 2**64, (1 << 127) | 5, 2**128 - 1]
    a = ids_to_array(vals)
    assert a.dtype == np.uint64 and a.shape == (6, 2) and ids_to_ints(a) == vals
    strs = ids_to_uuid_strs(a)
    assert strs[1] == "00000000-0000-0000-0000-000000000001" and ids_to_ints(ids_to_array(strs)) == vals
    assert ids_to_ints(ids_to_array([uuid.UUID(int=7)])) == [7]
    assert ids_to_ints(ids_to_array(np.array([3, 9], np.int64))) == [3, 9]
    # (hi, lo) lexicographic order == integer order == Postgres uuid byte order
    order = sorted(range(6), key=lambda i: (int(a[i, 0]), int(a[i, 1])))
    assert order == sorted(range(6), key=lambda i: vals[i])
    with pytest.raises(ValueError):
        ids_to_array([2**128])
    with pytest.raises(ValueError):
        ids_to_array(["not-a-uuid"])


def test_shard_function_is_balanced_and_stable():
    from outline_rag_b200.sharded import shard_of
    ids = O.ids_arange(0, 200_000)
    for world in (2, 4, 8):
        s = shard_of(ids, world)
        counts = np.bincount(s, minlength=world)
        assert counts.min() > 0.97 * 200_000 / world and counts.max() < 1.03 * 200_000 / world
        assert np.array_equal(s[:1000], shard_of(ids[:1000], world))        # depends on the id only
    wide = O.ids_from_ints([(i << 64) | 17 for i in range(4000)])            # only the high word varies
    assert np.bincount(shard_of(wide, 8), minlength=8).min() > 400


def test_doc_store_filter_grammar():
    from outline_rag_b200.vectorstore import MemoryDocStore
    st = MemoryDocStore()
    st.put_many(["a", "b", "c", "d"], ["t"] * 4,
                [{"source_id": "d1", "title": "x"}, {"source_id": "d1", "title": "y"},
                 {"source_id": "d2", "title": "x"}, {"source_id": "d3", "title": "z"}])
    assert st.ids_for_filter({"source_id": "d1"}) == ["a", "b"]
    assert st.ids_for_filter({"source_id": {"$in": ["d2", "d3"]}}) == ["c", "d"]
    assert st.ids_for_filter({"$and": [{"source_id": "d1"}, {"title": {"$eq": "y"}}]}) == ["b"]
    assert st.ids_for_filter({"$or": [{"title": "z"}, {"title": {"$ne": "x"}, "source_id": "d1"}]}) == ["b", "d"]
    assert st.ids_for_filter({"source_id": "nope"}) == []
    assert st.ids_for_source(["d1", "d3"]) == ["a", "b", "d"]
    with pytest.raises(NotImplementedError):
        st.ids_for_filter({"title": {"$like": "x%"}})


class FakeStoreIndex(FakeOwner):
    """Index stand-in that accepts whatever id forms `GpuVectorStore` passes (uuid strings)."""

    def upsert(self, ids, vecs):
        from outline_rag_b200.engine import ids_to_array
        super().upsert(ids_to_array(ids), vecs)

    def delete(self, ids):
        from outline_rag_b200.engine import ids_to_array
        return super().delete(ids_to_array(ids))

    def search_filtered(self, q, k, allow):
        from outline_rag_b200.engine import ids_to_array
        ok = {tuple(r) for r in ids_to_array(allow).tolist()}
        keep = np.array([tuple(r) in ok for r in self.ids.tolist()], bool)
        return FakeOwner(self.X[keep], self.ids[keep]).search(np.asarray(q, np.float32).reshape(1, -1), k)


class FakeEmb:
    def __init__(self, X):
        self.X = X

    async def aembed_documents(self, texts):
        return [self.X[int(t)] for t in texts]

    async def aembed_query(self, text):
        return self.X[int(text)]


def test_vectorstore_call_sequence_on_a_fake_index(small_table):
    """rag.py's order: look up the old chunk ids of the refreshed docs, adelete them, aadd_documents the
    re-chunked ones; ids default to uuid4; hits are hydrated in rank order; batching is transparent."""
    import outline_rag_b200 as orx
    X, _, _ = small_table
    owner = FakeStoreIndex(np.zeros((0, 1024), np.float32), np.zeros((0, 2), np.uint64))
    store = orx.GpuVectorStore(owner, FakeEmb(X), batch_window_ms=5.0)

    async def run():
        docs = [orx.Document(page_content=str(i), metadata={"source_id": f"d{i // 10}"}, id=str(uuid.UUID(int=i)))
                for i in range(100)]
        assert await store.aadd_documents(docs) == [d.id for d in docs]
        more = await store.aadd_documents([orx.Document(page_content="150", metadata={"source_id": "dx"})])
        assert len(more) == 1 and uuid.UUID(more[0]).version == 4
        hits = await asyncio.gather(*[store.as_retriever(search_kwargs={"k": 5}).ainvoke(str(i)) for i in (3, 42, 77)])
        assert [h[0].page_content for h in hits] == ["3", "42", "77"] and all(len(h) == 5 for h in hits)
        assert store.batcher.batches == 1                                    # three requests, one scan
        stale = store.doc_store.ids_for_source(["d4"])
        assert await store.adelete(ids=stale) is True and len(owner) == 91
        assert "42" not in [h.page_content for h in await store.asimilarity_search("42", k=12)]
        only_d7 = await asyncio.to_thread(store.similarity_search_with_score_by_vector, X[3], 12, {"source_id": "d7"})
        assert sorted(int(d.page_content) for d, _ in only_d7) == list(range(70, 80))
        assert [s for _, s in only_d7] == sorted(s for _, s in only_d7)
        await store.batcher.drain()

    asyncio.run(run())
    assert [op for op, _ in owner.log] == ["upsert", "upsert", "delete"]


def test_async_cold_start_coalesces_driver_chunks_and_always_frees_the_loader():
    """`aload_pgcopy`: psycopg yields tens of KB per chunk; the store hands the loader blocks of `feed_bytes`
    (one worker-thread hop each), closes it at the end, and closes it too when the stream breaks."""
    import outline_rag_b200 as orx

    class FakeLoader:
        def __init__(self):
            self.blocks, self.closed = [], 0

        def feed(self, data):
            self.blocks.append(bytes(data))

        def close(self):
            self.closed += 1
            return (sum(map(len, self.blocks)), 0)

    class FakeIndex:
        def __init__(self):
            self.loaders = []

        def pgcopy_loader(self):
            self.loaders.append(FakeLoader())
            return self.loaders[-1]

    payload = bytes(range(256)) * 400                                       # 102400 bytes

    async def chunks(fail_at=None):
        for n, i in enumerate(range(0, len(payload), 1000)):
            if fail_at is not None and n == fail_at:
                raise ConnectionError("server closed the connection")
            yield memoryview(payload)[i:i + 1000]

    ix = FakeIndex()
    store = orx.GpuVectorStore(ix, embedding_service=object())
    assert asyncio.run(store.aload_pgcopy(chunks(), feed_bytes=30_000)) == (len(payload), 0)
    ld = ix.loaders[0]
    assert b"".join(ld.blocks) == payload and [len(b) for b in ld.blocks] == [30_000, 30_000, 30_000, 12_400]
    assert ld.closed == 1
    with pytest.raises(ConnectionError):
        asyncio.run(store.aload_pgcopy(chunks(fail_at=50), feed_bytes=30_000))
    assert ix.loaders[1].closed == 1 and len(ix.loaders[1].blocks) == 1      # 30 KB were fed before the break
    assert "COPY (SELECT langchain_id, embedding FROM t " in store.COPY_SQL.format(table="t")


def test_sharded_index_hands_its_rank_to_the_loader():
    from outline_rag_b200.sharded import ShardedIndex

    class Local:
        def load_pgcopy(self, chunks, world, rank):
            self.seen = (chunks, world, rank)
            return (3, 1)

        def merge_topk(self, *a):
            raise AssertionError("not used")

    loc = Local()
    sh = ShardedIndex(local_index=loc)                                      # no process group: world 1, rank 0
    assert sh.load_pgcopy(b"stream") == (3, 1) and loc.seen == (b"stream", 1, 0)


def test_memory_doc_store_streams_its_embeddings_in_copy_binary_format():
    """The in-memory stand-in for `langchain_pg_embedding` emits what the COPY statement of the cold start
    streams: the oracle decodes it back to the stored rows, NULL embeddings included."""
    from oracle import pgvector_wire as W
    import outline_rag_b200 as orx
    st = orx.MemoryDocStore()
    rng = np.random.default_rng(1)
    X = rng.standard_normal((300, 1024)).astype(np.float32)
    ids = [str(uuid.UUID(int=3 * i + 1)) for i in range(300)]
    st.put_many(ids, ["c"] * 300, [{"source_id": "d"}] * 300, embeddings=X)
    st.put_many([str(uuid.UUID(int=2))], ["no vector yet"], [{}])                # NULL embedding
    st.delete_many(ids[10:20])
    chunks = list(st.copy_binary(rows_per_chunk=32))
    assert len(chunks) > 5 and max(map(len, chunks)) < 40 * 4200
    got_ids, got_X, n_null = W.copy_binary_parse(b"".join(chunks))
    keep = [i for i in range(300) if not 10 <= i < 20]
    assert n_null == 1 and O.ids_to_ints(got_ids) == [3 * i + 1 for i in keep]
    assert np.array_equal(got_X.view(np.uint32), X[keep].view(np.uint32))
    ld = orx.PgCopyLoader(None)                                                  # and the C walker agrees
    for c in chunks:
        ld.feed(c)
    assert ld.close() == (290, 1)


def test_prepared_filter_skips_the_doc_store_lookup(small_table):
    """`prepare_filter` resolves the metadata predicate once; later searches hand the device-resident handle
    straight to the index (no doc-store query, no id -> row resolution per call)."""
    import outline_rag_b200 as orx
    from outline_rag_b200.engine import Filter
    X, _, _ = small_table

    class Handle(Filter):                                   # a Filter without the C object behind it
        def __init__(self, ids):
            self.ids = ids

        def close(self):
            pass

    class Ix(FakeStoreIndex):
        def make_filter(self, ids):
            self.made = getattr(self, "made", 0) + 1
            return Handle(list(ids))

        def search_filtered(self, q, k, allow):
            self.last_allow = allow
            return super().search_filtered(q, k, allow.ids if isinstance(allow, Handle) else allow)

    owner = Ix(np.zeros((0, 1024), np.float32), np.zeros((0, 2), np.uint64))
    store = orx.GpuVectorStore(owner, FakeEmb(X))
    docs = [orx.Document(page_content=str(i), metadata={"source_id": f"d{i // 10}"}, id=str(uuid.UUID(int=i + 1)))
            for i in range(50)]
    asyncio.run(store.aadd_documents(docs))
    flt = store.prepare_filter({"source_id": {"$in": ["d1", "d3"]}})
    assert owner.made == 1 and len(flt.ids) == 20
    lookups = []
    real = store.doc_store.ids_for_filter
    store.doc_store.ids_for_filter = lambda f: lookups.append(f) or real(f)
    for probe in (12, 33):
        hits = store.similarity_search_with_score_by_vector(X[probe], 5, filter=flt)
        assert owner.last_allow is flt and hits[0][0].page_content == str(probe)
        assert all(d.metadata["source_id"] in ("d1", "d3") for d, _ in hits)
    assert lookups == []                                    # the prepared handle bypassed the doc store
    store.similarity_search_with_score_by_vector(X[12], 5, filter={"source_id": "d1"})
    assert lookups == [{"source_id": "d1"}]


def test_a_failed_batch_leaves_doc_store_and_index_untouched(small_table):
    """`aadd_documents` is all-or-nothing like the reference's INSERT: invalid vectors are rejected before any
    write, and when the durable write fails the device table is not touched either."""
    import outline_rag_b200 as orx
    X, _, _ = small_table

    class Emb(FakeEmb):
        async def aembed_documents(self, texts):
            return [self.bad if t == "bad" else self.X[int(t)] for t in texts]

    class FlakyStore(orx.MemoryDocStore):
        fail = False

        def put_many(self, *a, **kw):
            if self.fail:
                raise ConnectionError("server closed the connection unexpectedly")
            return super().put_many(*a, **kw)

    owner = FakeStoreIndex(np.zeros((0, 1024), np.float32), np.zeros((0, 2), np.uint64))
    emb, docs_store = Emb(X), FlakyStore()
    store = orx.GpuVectorStore(owner, emb, doc_store=docs_store)
    good = [orx.Document(page_content=str(i), id=str(uuid.UUID(int=i + 1))) for i in range(5)]
    asyncio.run(store.aadd_documents(good))
    for bad, msg in ((np.full(1024, np.nan, np.float32), "NaN or infinite"), (np.zeros(768, np.float32), "expected 1024 dimensions")):
        emb.bad = bad
        with pytest.raises(orx.OrxValueError, match=msg):
            asyncio.run(store.aadd_documents([orx.Document(page_content="7"), orx.Document(page_content="bad")]))
        assert len(owner) == 5 and len(docs_store._rows) == 5
    docs_store.fail = True
    with pytest.raises(ConnectionError):
        asyncio.run(store.aadd_documents([orx.Document(page_content="8")]))
    assert len(owner) == 5 and [op for op, _ in owner.log] == ["upsert"]
