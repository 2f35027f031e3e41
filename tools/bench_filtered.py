"""Filtered search (`orx_search_filtered`) at several selectivities: device time of the
bitmap scan (HBM traffic = eligible rows only) and wall time of the call, whose host part -- resolving
every allowed id to its row -- grows with the allow-list.

    python tools/bench_filtered.py [--rows 2000000] [--dtype fp32] > profiles/rN_filtered_scan.jsonl
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64, help="also time a batch of this many filtered queries per call")
    a = ap.parse_args()
    import torch
    import outline_rag_b200 as orx
    from bench import build_table
    from orx_testkit.synth import Synth, default_centres

    from bench import measured_peaks
    hbm_peak, _, _, peak_src = measured_peaks()
    syn = Synth(default_centres(a.rows))
    Q, _ = syn.queries(max(4, a.batch), min(a.rows, 100_000))
    elem = 4 if a.dtype == "fp32" else 2
    rng = np.random.default_rng(3)
    with orx.Index(a.dtype, capacity=a.rows) as ix:
        build_table(ix.upsert, 0, a.rows, 0, 1)
        from outline_rag_b200._lib import ORX_OPT_SCAN_TIMING
        ix.set_option(ORX_OPT_SCAN_TIMING, 1)          # `last_scan_ms` comes from event pairs around the scan launches
        ref = ix.search(Q[:1], 12)
        for frac in (1.0, 0.5, 0.1, 0.01):
            m = max(4096, int(a.rows * frac))
            sel = np.sort(rng.choice(a.rows, size=m, replace=False)) if m < a.rows else np.arange(a.rows)
            allow = np.zeros((m, 2), np.uint64)
            allow[:, 1] = sel.astype(np.uint64)
            scan_ms, wall_ms = [], []
            for it in range(a.iters + 2):
                t0 = time.perf_counter()
                got = ix.search_filtered(Q[:1], 12, allow)
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 2:
                    wall_ms.append(dt)
                    scan_ms.append(ix.stats()["last_scan_ms"])
            with ix.make_filter(allow) as flt:                      # resolved once, bitmap kept on the device
                handle_ms = []
                for it in range(a.iters + 2):
                    t0 = time.perf_counter()
                    got_h = ix.search_filtered(Q[:1], 12, flt)
                    dt = (time.perf_counter() - t0) * 1e3
                    if it >= 2:
                        handle_ms.append(dt)
                # a batch of filtered queries: one tcgen05 pass when batch x eligible >= rows, else one bitmap scan per query
                batch_ms = []
                for it in range(a.iters + 2):
                    t0 = time.perf_counter()
                    got_b = ix.search_filtered(Q[:a.batch], 12, flt)
                    dt = (time.perf_counter() - t0) * 1e3
                    if it >= 2:
                        batch_ms.append(dt)
                batch_path = ix.stats()["last_path"]
                batch_scan_ms = ix.stats()["last_scan_ms"]
                one = ix.search_filtered(Q[a.batch - 1:a.batch], 12, flt)          # the same query alone (bitmap GEMV scan)
                assert np.array_equal(one[0][0], got_b[0][-1]) and np.array_equal(one[1][0].view(np.uint64), got_b[1][-1].view(np.uint64))
            assert np.array_equal(got_h[0], got[0]) and np.array_equal(got_h[1].view(np.uint64), got[1].view(np.uint64))
            if frac == 1.0:
                assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1].view(np.uint64), ref[1].view(np.uint64))
            s = float(np.median(scan_ms))
            gbs = m * 1024 * elem / (s * 1e-3) / 1e9
            print(json.dumps({"rows": a.rows, "dtype": a.dtype, "eligible": m, "selectivity": m / a.rows,
                              "scan_ms": s, "scan_GBps_eligible_bytes": gbs,
                              "frac_of_hbm_peak": gbs / hbm_peak, "peak_source": peak_src,
                              "call_wall_ms": float(np.median(wall_ms)),
                              "call_wall_ms_with_filter_handle": float(np.median(handle_ms)),
                              "batch": a.batch, "batch_call_wall_ms": float(np.median(batch_ms)),
                              "batch_path": "tcgen05 (masked scale)" if batch_path == 2 else "bitmap GEMV per query",
                              "batch_last_scan_ms": batch_scan_ms,
                              "batch_speedup_vs_one_call_per_query": float(np.median(handle_ms)) * a.batch / float(np.median(batch_ms)),
                              "fallbacks": ix.stats()["fallback_exhaustive"]}), flush=True)


if __name__ == "__main__":
    main()
