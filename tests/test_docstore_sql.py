"""`SqlDocStore` (what stays in Postgres: content, metadata, the durable embedding column) on the reference's
table layout (app/database.py:118-131), driven through sqlite3 -- the standard library's DB-API 2.0 driver --
as the stand-in for psycopg: same SQL text up to the placeholder style and the `::vector` cast."""
import asyncio
import sqlite3
import uuid

import numpy as np
import pytest

import outline_rag_b200 as orx
from oracle import cosine_topk as O
from oracle import pgvector_wire as W
from tests.test_host_logic import FakeEmb, FakeStoreIndex

DDL = """
CREATE TABLE IF NOT EXISTS langchain_pg_embedding (
    langchain_id TEXT PRIMARY KEY,
    content TEXT,
    embedding TEXT,
    source_id TEXT,
    title TEXT,
    outline_updated_at_str TEXT,
    url TEXT
);
CREATE INDEX IF NOT EXISTS idx_langchain_embedding_source_id ON langchain_pg_embedding(source_id);
"""


@pytest.fixture()
def store(tmp_path):
    path = str(tmp_path / "rag.db")
    conn = sqlite3.connect(path)
    conn.executescript(DDL)
    conn.close()
    return orx.SqlDocStore(lambda: sqlite3.connect(path), paramstyle="qmark")


def test_vector_text_round_trips_bit_for_bit():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1024) * 10.0 ** rng.integers(-8, 8, size=1024)).astype(np.float32)
    x[:4] = [0.0, -0.0, np.float32(1e-42), np.finfo(np.float32).max]
    text = orx.vector_to_text(x)
    assert text.startswith("[") and " " not in text
    assert np.array_equal(orx.parse_vector_text(text).view(np.uint32), x.view(np.uint32))
    assert np.array_equal(W.vector_in(text).view(np.uint32), x.view(np.uint32))


def test_contract_put_get_delete_lookup(store):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((6, 1024)).astype(np.float32)
    ids = [str(uuid.UUID(int=i + 1)) for i in range(6)]
    metas = [{"source_id": f"d{i // 2}", "title": f"t{i}", "outline_updated_at_str": "2026", "url": f"/u{i}"} for i in range(6)]
    store.put_many(ids, [f"c{i}" for i in range(6)], metas, embeddings=X)
    store.put_many(ids[:1], ["c0 v2"], [dict(metas[0], title="new")], embeddings=X[5:6])         # ON CONFLICT DO UPDATE
    got = store.get_many([ids[3], "00000000-0000-0000-0000-0000000000ff", ids[0]])
    assert got[0] == ("c3", metas[3]) and got[1] is None and got[2] == ("c0 v2", dict(metas[0], title="new"))
    assert sorted(store.ids_for_source(["d1", "d2"])) == ids[2:6] and store.ids_for_source([]) == []
    assert store.ids_for_filter({"source_id": "d0"}) == ids[:2]
    assert sorted(store.ids_for_filter({"$or": [{"title": "t5"}, {"$and": [{"source_id": {"$in": ["d1"]}}, {"url": {"$ne": "/u2"}}]}]})) \
        == [ids[3], ids[5]]
    assert store.ids_for_filter({"source_id": {"$in": []}}) == []
    with pytest.raises(ValueError, match="not a metadata column"):
        store.ids_for_filter({"content; DROP TABLE x": 1})
    with pytest.raises(NotImplementedError):
        store.ids_for_filter({"title": {"$like": "t%"}})
    store.delete_many(ids[4:])
    assert store.get_many(ids[4:]) == [None, None]
    # the durable embeddings come back as the COPY BINARY stream of the cold start, bit for bit
    s_ids, s_X, n_null = W.copy_binary_parse(b"".join(store.copy_binary(rows_per_chunk=3)))
    assert n_null == 0 and O.ids_to_ints(s_ids) == [1, 2, 3, 4]
    assert np.array_equal(s_X.view(np.uint32), np.concatenate([X[5:6], X[1:4]]).view(np.uint32))
    with pytest.raises(ValueError, match="identifier"):
        orx.SqlDocStore(lambda: None, table="t; drop")


def test_vector_store_runs_the_reference_sequence_on_the_sql_doc_store(store, small_table):
    """rag.py's refresh order against the SQL-backed doc store: ids by source_id -> adelete -> aadd_documents;
    hits are hydrated from SQL in rank order; the metadata filter is resolved by SQL."""
    X, _, _ = small_table
    owner = FakeStoreIndex(np.zeros((0, 1024), np.float32), np.zeros((0, 2), np.uint64))
    vs = orx.GpuVectorStore(owner, FakeEmb(X), doc_store=store)

    async def run():
        docs = [orx.Document(page_content=str(i), metadata={"source_id": f"d{i // 10}", "title": f"T{i}"},
                             id=str(uuid.UUID(int=i + 1))) for i in range(60)]
        await vs.aadd_documents(docs)
        hits = await vs.asimilarity_search("17", k=4)
        assert hits[0].page_content == "17" and hits[0].metadata["source_id"] == "d1" and hits[0].metadata["title"] == "T17"
        stale = store.ids_for_source(["d1"])
        assert len(stale) == 10 and await vs.adelete(ids=stale) is True
        assert "17" not in [h.page_content for h in await vs.asimilarity_search("17", k=12)]
        only = await asyncio.to_thread(vs.similarity_search_with_score_by_vector, X[33], 12, {"source_id": "d3"})
        assert sorted(int(d.page_content) for d, _ in only) == list(range(30, 40))

    asyncio.run(run())
    assert len(list(store.copy_binary())) >= 3 and len(owner) == 50
