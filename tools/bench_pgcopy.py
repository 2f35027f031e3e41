"""Cold-start load rate: a COPY BINARY stream of (langchain_id, embedding) in host memory -> device table
(`orx_pgcopy_*`), beside the NumPy decode of the same stream on the host cores.

    python tools/bench_pgcopy.py [--rows 262144] [--dtype fp32] [--chunk 65536] > profiles/rN_pgcopy_load.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def make_stream(n: int, seed: int = 7, first: int = 1) -> bytes:
    from outline_rag_b200.pgwire import encode_copy_binary
    rng = np.random.default_rng(seed)
    ids = np.zeros((n, 2), np.uint64)
    ids[:, 1] = np.arange(first, first + n, dtype=np.uint64)
    return encode_copy_binary(ids, rng.standard_normal((n, 1024), dtype=np.float32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=262144)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--chunk", type=int, default=65536, help="bytes per feed call (psycopg yields tens of KB)")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--streams", type=int, default=1,
                    help="parallel COPY streams into ONE index (N connections, each `WHERE` on a slice of the ids), one thread each")
    ap.add_argument("--lib", default=None, help="A/B: load this build of liborx.so instead of the in-tree default")
    a = ap.parse_args()
    if a.lib:
        import types
        sys.modules["orx_lib_override"] = types.SimpleNamespace(LIB_PATH=os.path.abspath(a.lib))
    import torch
    import outline_rag_b200 as orx

    import threading
    per = a.rows // a.streams
    a.rows = per * a.streams
    streams = [make_stream(per, seed=7 + i, first=1 + i * per) for i in range(a.streams)]
    stream = streams[0]
    views = [memoryview(b) for b in streams]
    best = None
    for _ in range(a.repeat):
        with orx.Index(a.dtype, capacity=a.rows) as ix:
            results = [None] * a.streams

            def feed_one(i):
                with ix.pgcopy_loader() as ld:
                    mv = views[i]
                    for o in range(0, len(mv), a.chunk):
                        ld.feed(mv[o:o + a.chunk])
                results[i] = ld.result

            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if a.streams == 1:
                feed_one(0)
            else:
                th = [threading.Thread(target=feed_one, args=(i,)) for i in range(a.streams)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert results == [(per, 0)] * a.streams and len(ix) == a.rows, (results, len(ix))
            launches = ix.stats()["kernel_launches"]
            probe = np.array([0, per // 2, per - 1])
            ids = np.zeros((3, 2), np.uint64)
            ids[:, 1] = probe + 1
            got, found = ix.fetch(ids)
        best = dt if best is None else min(best, dt)
    # the same decode on the host: framing is fixed-stride here, so NumPy's byte swap is the whole job
    t0 = time.perf_counter()
    from outline_rag_b200.pgwire import tuple_dtype
    t = np.frombuffer(stream, tuple_dtype(1024), count=per, offset=19)
    X = t["v"].astype(np.float32)
    ids_h = t["id"].astype(np.uint64)
    cpu = time.perf_counter() - t0
    if a.dtype == "fp32":
        assert found.all() and np.array_equal(got.view(np.uint32), X[probe].view(np.uint32))
    print(json.dumps({
        "metric": "cold_start_load_rows_per_s", "value": a.rows / best, "unit": "rows/s", "rows": a.rows, "dtype": a.dtype,
        "streams": a.streams, "stream_GB": len(stream) * a.streams / 1e9,
        "stream_GB_per_s": len(stream) * a.streams / 1e9 / best, "seconds": best,
        "feed_chunk_bytes": a.chunk, "lib": a.lib or "in-tree", "gpu_launches": int(launches),
        "cpu_baseline": {"kind": "port", "what": "NumPy frombuffer + big-endian -> fp32 astype of the same stream (decode only, no table build)",
                         "seconds": cpu, "rows_per_s": per / cpu, "cores": 1, "rows": per},
        "note": "e2e: host stream -> pinned staging -> H2D -> decode kernel -> validate + commit (norms, id map)"}))


if __name__ == "__main__":
    main()
