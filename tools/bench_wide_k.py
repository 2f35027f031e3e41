"""Batches with a wider k (a larger reranker feed): wall time per `Index.search(Q[batch], k)` call for
k = 12 ... 128, the scan path taken, and a full oracle check of a few queries of every call.

    python tools/bench_wide_k.py [--rows 2000000] [--dtype fp32] [--batch 64] > profiles/rN_wide_k.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--verify", type=int, default=2, help="queries per call checked against a full oracle pass")
    ap.add_argument("--ks", default="12,32,33,64,100,128")
    ap.add_argument("--lib", default=None, help="A/B: load this build of liborx.so instead of the in-tree default")
    a = ap.parse_args()
    if a.lib:
        import types
        sys.modules["orx_lib_override"] = types.SimpleNamespace(LIB_PATH=os.path.abspath(a.lib))
    import outline_rag_b200 as orx
    from bench import build_table
    from oracle import cosine_topk as O
    from orx_testkit.synth import Synth, default_centres

    syn = Synth(default_centres(a.rows))
    Q, _ = syn.queries(a.batch, min(a.rows, 100_000))
    with orx.Index(a.dtype, capacity=a.rows) as ix:
        build_table(ix.upsert, 0, a.rows, 0, 1)
        ids = np.zeros((a.rows, 2), np.uint64)
        rows = np.zeros((a.rows, 1024 * (4 if a.dtype == "fp32" else 2)), np.uint8)
        if a.verify:
            ix.export_rows(0, a.rows, ids, rows)
            X = rows.view(np.float32) if a.dtype == "fp32" else O.StreamingTopK.bf16_bits_to_f32(rows.view(np.uint16))
        for k in [int(v) for v in a.ks.split(",")]:
            ms = []
            for it in range(a.iters + 2):
                t0 = time.perf_counter()
                got = ix.search(Q, k)
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 2:
                    ms.append(dt)
            st = ix.stats()
            ok = True
            for i in range(a.verify):
                w_ids, w_d = O.topk_exact(X, ids, Q[i], k)
                ok = ok and np.array_equal(got[0][i], w_ids) and np.array_equal(got[1][i].view(np.uint64), w_d.view(np.uint64))
            one = []
            for it in range(3):
                t0 = time.perf_counter()
                ix.search(Q[:1], k)
                one.append((time.perf_counter() - t0) * 1e3)
            print(json.dumps({"rows": a.rows, "dtype": a.dtype, "batch": a.batch, "k": k, "lib": a.lib or "in-tree",
                              "call_wall_ms": float(np.median(ms)), "last_scan_ms": st["last_scan_ms"],
                              "path": "tcgen05" if st["last_path"] == 2 else "gemv per query",
                              "single_query_call_ms": float(np.median(one)),
                              "speedup_vs_one_pass_per_query": float(np.median(one)) * a.batch / float(np.median(ms)),
                              "fallback_gemv": st["fallback_gemv"], "fallback_exhaustive": st["fallback_exhaustive"],
                              "oracle_full_scan_equal": bool(ok), "queries_checked": a.verify}), flush=True)


if __name__ == "__main__":
    main()
