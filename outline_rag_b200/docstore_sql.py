"""`SqlDocStore`: the doc-store side of `GpuVectorStore` on a SQL table, through any DB-API 2.0 connection.

What stays in Postgres when the similarity search moves to the GPU: the table
`langchain_pg_embedding` of the reference (app/database.py:118-131) with `content`, the four metadata columns
(rag.py:73-78) and -- still written, for durability and cold start -- the `embedding` column.  The raw SQL the
reference itself issues against that table keeps working unchanged (rag.py:216-224, 278-286, 357-363).

The store speaks plain DB-API 2.0 (PEP 249), so it runs on psycopg 3 connections in production
(`paramstyle="format"`, `embedding_cast="::vector"`) and on the standard library's sqlite3 in the CPU tests
(`paramstyle="qmark"`); `GpuVectorStore` calls it from worker threads (`asyncio.to_thread`), never on the event loop.
`connect` is a zero-argument callable returning a connection; one connection is opened per call and closed after
it, which is what a pool's `getconn` / context manager gives.

Only whitelisted column names ever reach the SQL text; every value travels as a bound parameter.
"""
from __future__ import annotations

import uuid
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from .engine import parse_vector_text
from .pgwire import COPY_HEADER, COPY_TRAILER, encode_tuples

DEFAULT_METADATA_COLUMNS = ["source_id", "title", "outline_updated_at_str", "url"]   # reference rag.py:73-78


def vector_to_text(x) -> str:
    """`'[v0,v1,...]'` with the shortest decimal that round-trips each fp32 value (what pgvector's
    `vector_out` prints, and what `vector_in` / `parse_vector_text` read back bit for bit)."""
    a = np.asarray(x, dtype=np.float32)
    return "[" + ",".join(map(str, a)) + "]"          # str(np.float32) is the shortest round-trip decimal


class SqlDocStore:
    stores_embeddings = True

    def __init__(self, connect: Callable[[], object], table: str = "langchain_pg_embedding",
                 metadata_columns: Optional[Sequence[str]] = None, paramstyle: str = "format",
                 embedding_cast: str = "", id_column: str = "langchain_id", content_column: str = "content",
                 embedding_column: str = "embedding"):
        self.connect = connect
        self.metadata_columns = list(metadata_columns or DEFAULT_METADATA_COLUMNS)
        for name in [table, id_column, content_column, embedding_column, *self.metadata_columns]:
            if not name.replace("_", "").isalnum():
                raise ValueError(f"not a plain SQL identifier: {name!r}")
        if paramstyle not in ("format", "qmark"):
            raise ValueError("paramstyle must be 'format' (%s: psycopg) or 'qmark' (?: sqlite3)")
        self.table, self.id_col, self.content_col, self.emb_col = table, id_column, content_column, embedding_column
        self.ph = "%s" if paramstyle == "format" else "?"
        self.embedding_cast = embedding_cast            # "::vector" on Postgres: the parameter is pgvector's text form

    # ------------------------------------------------------------------ plumbing
    def _run(self, sql: str, params: Sequence = (), many: bool = False, fetch: bool = False):
        conn = self.connect()
        try:
            cur = conn.cursor()
            if many:
                cur.executemany(sql, params)
            else:
                cur.execute(sql, tuple(params))
            rows = cur.fetchall() if fetch else None
            conn.commit()
            return rows
        finally:
            conn.close()

    def _in(self, n: int) -> str:
        return "(" + ",".join([self.ph] * n) + ")"

    # ------------------------------------------------------------------ the DocStore contract
    def put_many(self, ids: Sequence[str], contents: Sequence[str], metadatas: Sequence[dict], embeddings=None) -> None:
        """`INSERT ... ON CONFLICT (langchain_id) DO UPDATE` per row -- the statement langchain-postgres issues
        for `aadd_documents` (reference rag.py:235), embedding included so Postgres stays the source of truth."""
        cols = [self.id_col, self.content_col, *self.metadata_columns, self.emb_col]
        values = [self.ph] * (len(cols) - 1) + [self.ph + self.embedding_cast]
        sets = ", ".join(f"{c} = excluded.{c}" for c in cols[1:])
        sql = (f"INSERT INTO {self.table} ({', '.join(cols)}) VALUES ({', '.join(values)}) "
               f"ON CONFLICT ({self.id_col}) DO UPDATE SET {sets}")
        rows = []
        for n, (i, c, m) in enumerate(zip(ids, contents, metadatas)):
            emb = vector_to_text(embeddings[n]) if embeddings is not None else None
            rows.append((str(i), c, *[m.get(k) for k in self.metadata_columns], emb))
        if rows:
            self._run(sql, rows, many=True)

    def get_many(self, ids: Sequence[str]) -> list[Optional[tuple[str, dict]]]:
        """`SELECT content, <metadata> ... WHERE langchain_id IN (...)`, returned in the order of `ids`."""
        ids = [str(i) for i in ids]
        if not ids:
            return []
        cols = [self.id_col, self.content_col, *self.metadata_columns]
        got = self._run(f"SELECT {', '.join(cols)} FROM {self.table} WHERE {self.id_col} IN {self._in(len(ids))}", ids,
                        fetch=True)
        by_id = {str(r[0]): (r[1], dict(zip(self.metadata_columns, r[2:]))) for r in got}
        return [by_id.get(i) for i in ids]

    def delete_many(self, ids: Iterable[str]) -> None:
        ids = [str(i) for i in ids]
        if ids:
            self._run(f"DELETE FROM {self.table} WHERE {self.id_col} IN {self._in(len(ids))}", ids)

    def ids_for_source(self, source_ids: Iterable[str]) -> list[str]:
        """reference app/rag.py:216-224: the chunk ids of the documents being refreshed."""
        src = list(source_ids)
        if not src:
            return []
        got = self._run(f"SELECT {self.id_col} FROM {self.table} WHERE source_id IN {self._in(len(src))}", src, fetch=True)
        return [str(r[0]) for r in got]

    def ids_for_filter(self, flt: dict) -> list[str]:
        """The upstream filter grammar subset of `MemoryDocStore.ids_for_filter`, compiled to a WHERE clause."""
        where, params = self._where(flt)
        got = self._run(f"SELECT {self.id_col} FROM {self.table} WHERE {where}", params, fetch=True)
        return [str(r[0]) for r in got]

    def _where(self, f: dict) -> tuple[str, list]:
        parts, params = [], []
        for key, cond in f.items():
            if key in ("$and", "$or"):
                subs = [self._where(c) for c in cond]
                if not subs:
                    parts.append("1=1" if key == "$and" else "1=0")
                    continue
                parts.append("(" + (" AND " if key == "$and" else " OR ").join(f"({w})" for w, _ in subs) + ")")
                for _, p in subs:
                    params += p
                continue
            if key not in self.metadata_columns:
                raise ValueError(f"not a metadata column: {key!r}")
            if not isinstance(cond, dict):
                cond = {"$eq": cond}
            for op, arg in cond.items():
                if op == "$eq":
                    parts.append(f"{key} = {self.ph}")
                    params.append(arg)
                elif op == "$ne":                    # NULL-safe, like Python's `!=` in MemoryDocStore
                    parts.append(f"({key} IS NULL OR {key} <> {self.ph})")
                    params.append(arg)
                elif op == "$in":
                    arg = list(arg)
                    parts.append(f"{key} IN {self._in(len(arg))}" if arg else "1=0")
                    params += arg
                else:
                    raise NotImplementedError(f"filter operator {op}")
        return (" AND ".join(parts) if parts else "1=1"), params

    # ------------------------------------------------------------------ cold start
    def copy_binary(self, rows_per_chunk: int = 1024):
        """The `COPY (SELECT langchain_id, embedding ...) TO STDOUT (FORMAT binary)` stream of this table, for
        `GpuVectorStore.load_pgcopy`.  Generic DB-API path (SELECT the text form, re-encode); on Postgres use the
        server's own COPY through psycopg's `cursor.copy(GpuVectorStore.COPY_SQL...)` instead -- it skips the text."""
        got = self._run(f"SELECT {self.id_col}, {self.emb_col} FROM {self.table} WHERE {self.emb_col} IS NOT NULL", fetch=True)
        yield COPY_HEADER
        for s in range(0, len(got), rows_per_chunk):
            part = got[s:s + rows_per_chunk]
            ids = [uuid.UUID(str(r[0])).int for r in part]
            vecs = [parse_vector_text(str(r[1]), dim=str(r[1]).count(",") + 1) for r in part]
            yield encode_tuples(ids, np.stack(vecs))
        yield COPY_TRAILER


__all__ = ["SqlDocStore", "vector_to_text"]
