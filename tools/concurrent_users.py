"""Concurrent-user scenario (reference app/blueprints/api.py:122 issues ONE retrieval per request):
C asyncio clients each run R sequential by-vector searches through `GpuVectorStore`, with and without
the micro-batching front end.  Prints one JSON line per (C, window) with QPS and latency percentiles.

    python tools/concurrent_users.py --rows 10000000 --dtype fp32
"""
import argparse
import asyncio
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import outline_rag_b200 as orx                                   # noqa: E402
from bench import build_table                                    # noqa: E402
from orx_testkit.synth import Synth, default_centres        # noqa: E402


class NoEmb:                                                     # queries arrive as vectors
    pass


class AnyDocStore(orx.MemoryDocStore):
    """Stands in for Postgres holding content + metadata of all 10M chunks: every id has a row."""

    def get_many(self, ids):
        return [("chunk " + i, {"source_id": "doc"}) for i in ids]


async def scenario(store, Q, clients, rounds):
    lat = []

    async def user(c):
        for r in range(rounds):
            q = Q[(c * rounds + r) % Q.shape[0]]
            t = time.perf_counter()
            res = await store.asimilarity_search_with_score_by_vector(q, k=orx.TOP_K)
            lat.append(time.perf_counter() - t)
            assert len(res) == orx.TOP_K
    t0 = time.perf_counter()
    await asyncio.gather(*[user(c) for c in range(clients)])
    wall = time.perf_counter() - t0
    if store.batcher is not None:
        await store.batcher.drain()
    return clients * rounds / wall, np.asarray(lat)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--clients", type=int, nargs="+", default=[1, 16, 64, 256])
    a = ap.parse_args()
    ix = orx.Index(a.dtype, a.rows, 0)
    build_table(ix.upsert, 0, a.rows, 0, 1)
    Q, _ = Synth(default_centres(a.rows)).queries(512, a.rows)
    for window in (None, 1.0):
        store = orx.GpuVectorStore(ix, NoEmb(), doc_store=AnyDocStore(), batch_window_ms=window, max_batch=256)
        for c in a.clients:
            rounds = max(4, min(40, 1024 // c))
            asyncio.run(scenario(store, Q, min(c, 8), 2))                        # warm-up
            s0 = ix.stats()
            qps, lat = asyncio.run(scenario(store, Q, c, rounds))
            s1 = ix.stats()
            print(json.dumps({"rows": a.rows, "dtype": a.dtype, "clients": c, "batch_window_ms": window,
                              "requests": c * rounds, "qps": round(qps, 1),
                              "p50_ms": round(float(np.median(lat)) * 1e3, 3),
                              "p99_ms": round(float(np.percentile(lat, 99)) * 1e3, 3),
                              "scans": int(s1["searches"] - s0["searches"])}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
