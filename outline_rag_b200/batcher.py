"""Query micro-batching front end.

The reference issues ONE retrieval per `/chat/api/ask` request (reference app/blueprints/api.py:122 ->
`compression_retriever.ainvoke(query)`), so concurrent users turn into many independent
memory-bound scans.  `QueryBatcher` coalesces the by-vector searches that arrive within a short
window into ONE batched `Index.search` call: the table is then read once for the whole batch
(tcgen05 scan, `csrc/scan_umma.cu`) instead of once per request.

Semantics are unchanged: every caller gets exactly the result its own `search(q, k)` would have
returned -- the ordering (distance ASC, NaN last, id ASC) is total, so the top-k' of a query is a
prefix of its top-k and requests with different k share a batch run at max(k).  Requests are grouped
by what one engine call can serve: the same prepared `Filter` (or none), and the same k class -- the
engine's batched scan keeps 64-key candidate lists up to k = 16, 160-key lists up to k = 64 and takes
the exact fp32 path beyond (csrc/scan_umma.cu), so one wide request must not drag a window of k = 12
requests onto a slower path.  An error that
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
        self._pending: list[tuple[np.ndarray, int, asyncio.Future, object]] = []
        self._timer: Optional[asyncio.TimerHandle] = None
        self._inflight: set[asyncio.Task] = set()
        self.batches = 0           # searches actually issued
        self.requests = 0          # requests answered

    K_CLASSES = (16, 64)           # upper bounds of the k classes served together (beyond the last: its own class)

    @classmethod
    def _k_class(cls, k: int) -> int:
        for c, top in enumerate(cls.K_CLASSES):
            if k <= top:
                return c
        return len(cls.K_CLASSES)

    async def search(self, embedding, k: int = 12, filter=None):
        """-> (ids uint64 [m, 2], distance float64 [m]) with m = min(k, live rows).  ``filter``: a prepared
        `engine.Filter` handle; requests carrying the SAME handle are answered by one `search_filtered` call."""
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
        self._pending.append((q, int(k), fut, filter))
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
        groups: dict[tuple[int, int], list] = {}
        for item in batch:
            groups.setdefault((id(item[3]) if item[3] is not None else 0, self._k_class(item[1])), []).append(item)
        for group in groups.values():
            await self._run_group(group)

    async def _run_group(self, group) -> None:
        Q = np.stack([b[0] for b in group])
        kmax = max(b[1] for b in group)
        flt = group[0][3]
        try:
            if flt is None:
                ids, dist, cnt = await asyncio.to_thread(self.index.search, Q, kmax)
            else:
                ids, dist, cnt = await asyncio.to_thread(self.index.search_filtered, Q, kmax, flt)
        except Exception as e:                             # engine failure: every waiter of this call sees it
            for _, _, fut, _ in group:
                if not fut.done():
                    fut.set_exception(e)
            return
        self.batches += 1
        self.requests += len(group)
        for i, (_, k, fut, _) in enumerate(group):
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
