"""outline_rag_b200 -- B200-native exact cosine top-k retrieval for Outline-RAG's hot path.

Stands in for the pgvector query behind ``rag.vector_store`` (reference app/rag.py:69-87):
``ORDER BY embedding <=> :q LIMIT TOP_K`` plus its upsert/delete path.  Importing this
package loads ``liborx.so`` (hand-written sm_100a CUDA behind the C-ABI of
``include/orx.h``); there is no CPU fallback -- a missing library is an ImportError.
"""
from ._lib import (DTYPE_BF16, DTYPE_F32, ORX_DIM, ORX_MAX_K, OrxError, OrxValueError, build, version)
from .batcher import QueryBatcher
from .daemon import IndexServer, RemoteIndex, serve_in_thread
from .engine import Filter, Index, PgCopyLoader, parse_vector_text, ids_to_array, ids_to_ints, ids_to_uuid_strs
from . import pgwire
from .docstore_sql import SqlDocStore, vector_to_text
from .pgwire import encode_copy_binary
from .vectorstore import Document, GpuRetriever, GpuVectorStore, MemoryDocStore

TOP_K = 12               # reference app/config.py:253
REFRESH_BATCH_SIZE = 50  # README.md:42 (code default 100, app/config.py:255); BASELINE.json uses 50

__all__ = ["Index", "Filter", "PgCopyLoader", "encode_copy_binary", "SqlDocStore", "vector_to_text", "parse_vector_text", "QueryBatcher", "IndexServer", "RemoteIndex", "serve_in_thread", "GpuVectorStore", "GpuRetriever", "MemoryDocStore", "Document", "OrxError",
           "OrxValueError", "ids_to_array", "ids_to_ints", "ids_to_uuid_strs",
           "build", "version", "ORX_DIM", "ORX_MAX_K", "DTYPE_F32", "DTYPE_BF16", "TOP_K",
           "REFRESH_BATCH_SIZE"]
