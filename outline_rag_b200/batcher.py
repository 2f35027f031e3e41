"""Query micro-batching front end.

The reference issues ONE retrieval per `/chat/api/ask` request (reference app/blueprints/api.py:122 ->
`compression_retriever.ainvoke(query)`), so concurrent users turn into many independent
memory-bound scans.  `QueryBatcher` coalesces the by-vector searches that arrive within a short
window into ONE batched `Index.search` call: the table is then read once for the whole batch
(tcgen05 scan, `csrc/scan_umma.cu`) instead of once per request.

Semantics are unchanged: every caller gets exactly the result its own `search(q, k)` would have
returned -- the ordering (distance ASC, NaN last, id ASC) is total, so the top-k' of a query is a
prefix of its top-k and requests with different k share a batch run at max(k).  An error that
concerns one request only (wrong dimension, NaN/Inf in the query; pgvector rejects those per
statement) fails that request alone: the offending rows are screened out before the batch is
submitted.  Pure host logic; the arithmetic stays in the C-ABI.
"""
from __future__ import annotations

import asyncio
from typing import Optional

import numpy as np

from ._lib import ORX_DIM, ORX_ERR_DIM, ORX_ERR_NONFINITE, ORX_MAX_K, OrxValueError


class QueryBatcher:
    def __init__(self, index, max_batch: int = 256, max_wait_ms: float = 1.0):
        self.index = index
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._pending: list[tuple[np.ndarray, int, asyncio.Future]] = []
        self._timer: Optional[asyncio.TimerHandle] = None
        self._inflight: set[asyncio.Task] = set()
        self.batches = 0           # searches actually issued
        self.requests = 0          # requests answered

    async def search(self, embedding, k: int = 12):
        """-> (ids uint64 [m, 2], distance float64 [m]) with m = min(k, live rows)."""
        loop = asyncio.get_running_loop()
        q = np.asarray(embedding, dtype=np.float32).reshape(-1)
        fut: asyncio.Future = loop.create_future()
        # per-request validation (the same errors orx_search raises), so one bad query cannot fail a batch
        if q.shape[0] != ORX_DIM:
            fut.set_exception(OrxValueError(ORX_ERR_DIM, f"different vector dimensions {ORX_DIM} and {q.shape[0]}"))
            return await fut
        if not np.isfinite(q).all():
            fut.set_exception(OrxValueError(ORX_ERR_NONFINITE, "NaN or infinite value not allowed in vector"))
            return await fut
        if not (1 <= int(k) <= ORX_MAX_K):
            fut.set_exception(OrxValueError(-1, f"k must be in [1, {ORX_MAX_K}], got {k}"))
            return await fut
        self._pending.append((q, int(k), fut))
        if len(self._pending) >= self.max_batch:
            self._flush()
        elif self._timer is None:
            self._timer = loop.call_later(self.max_wait, self._flush)
        return await fut

    def _flush(self) -> None:
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None
        batch, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
        if not batch:
            return
        task = asyncio.get_running_loop().create_task(self._run(batch))
        self._inflight.add(task)
        task.add_done_callback(self._inflight.discard)
        if self._pending:                                  # overflow of a burst: next window immediately
            self._timer = asyncio.get_running_loop().call_later(0, self._flush)

    async def _run(self, batch) -> None:
        Q = np.stack([b[0] for b in batch])
        kmax = max(b[1] for b in batch)
        try:
            ids, dist, cnt = await asyncio.to_thread(self.index.search, Q, kmax)
        except Exception as e:                             # engine failure: every waiter sees it
            for _, _, fut in batch:
                if not fut.done():
                    fut.set_exception(e)
            return
        self.batches += 1
        self.requests += len(batch)
        for i, (_, k, fut) in enumerate(batch):
            if fut.done():
                continue
            m = min(int(cnt[i]), k)
            fut.set_result((ids[i, :m].copy(), dist[i, :m].copy()))

    async def drain(self) -> None:
        """Flush what is pending and wait for the in-flight batches (shutdown / tests)."""
        self._flush()
        while self._inflight:
            await asyncio.gather(*list(self._inflight), return_exceptions=True)


__all__ = ["QueryBatcher"]
