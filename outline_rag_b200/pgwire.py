"""Host-side encoders for the PostgreSQL / pgvector wire formats (the decoders are in `csrc/pgwire.cu`).

`encode_copy_binary(ids, vecs)` produces the byte stream of `COPY ... (FORMAT binary)` for rows of
`(langchain_id uuid, embedding vector(dim))` -- what `GpuVectorStore.COPY_SQL` reads back at cold start and
what a maintainer can send with psycopg 3's `cursor.copy("COPY staging (langchain_id, embedding) FROM STDIN
(FORMAT binary)")` to keep the durable copy of the embeddings in Postgres (reference table: app/database.py:
118-131) without formatting 1024 floats per row as text, which is what the reference's
`aadd_documents` path does today (langchain-postgres sends `str(list)`, rag.py:235).
Pure NumPy byte shuffling; no arithmetic.
"""
from __future__ import annotations

import numpy as np

from .engine import ids_to_array

COPY_SIGNATURE = b"PGCOPY\n\xff\r\n\x00"
COPY_HEADER = COPY_SIGNATURE + bytes(8)          # flags = 0, no header extension
COPY_TRAILER = b"\xff\xff"                       # int16 -1


def tuple_dtype(dim: int) -> np.dtype:
    """One COPY BINARY tuple of (uuid, vector(dim)) as a packed big-endian record."""
    return np.dtype([("nf", ">i2"), ("l1", ">i4"), ("id", ">u8", (2,)), ("l2", ">i4"), ("dim", ">i2"),
                     ("unused", ">i2"), ("v", ">f4", (dim,))])


def encode_tuples(ids, vecs) -> bytes:
    """The tuples only (no header / trailer), for streaming a large table chunk by chunk."""
    ida = ids_to_array(ids)
    X = np.ascontiguousarray(vecs, dtype=np.float32)
    if X.ndim != 2 or X.shape[0] != ida.shape[0]:
        raise ValueError(f"{ida.shape[0]} ids for vectors of shape {X.shape}")
    dim = X.shape[1]
    if not 1 <= dim <= 16000:
        raise ValueError("vector must have between 1 and 16000 dimensions")
    t = np.empty(X.shape[0], tuple_dtype(dim))
    t["nf"], t["l1"], t["l2"], t["dim"], t["unused"] = 2, 16, 4 + 4 * dim, dim, 0
    t["id"] = ida
    t["v"] = X
    return t.tobytes()


def encode_copy_binary(ids, vecs) -> bytes:
    """A complete COPY BINARY stream (header, tuples, end-of-data marker)."""
    return COPY_HEADER + encode_tuples(ids, vecs) + COPY_TRAILER


__all__ = ["encode_copy_binary", "encode_tuples", "tuple_dtype", "COPY_HEADER", "COPY_TRAILER"]
