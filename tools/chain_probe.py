#!/usr/bin/env python
"""Where does the time above the scan kernel go?  Calls the C-ABI directly (pre-allocated buffers, no Python
allocation inside the loop) and prints, per variant, the closed-loop latency of one single-query search and what is
left after the scan kernel's own device time.  Run under gpurun; pair it with an ncu launch list of the same command
for the per-kernel durations (prep / scan / finalize)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--devices", default="")          # e.g. "0,0" or "0,1,2,3": the multi-GPU index in one process
    a = ap.parse_args()
    import torch
    import outline_rag_b200 as orx
    from outline_rag_b200._lib import lib, check, ORX_OPT_SCAN_TIMING
    from orx_testkit.device import synth_rows_device
    from orx_testkit.synth import SEED_TABLE, Synth, default_centres
    devs = [int(x) for x in a.devices.split(",")] if a.devices else None
    ix = orx.Index(a.dtype, a.rows, 0, devices=devs)
    nc = default_centres(a.rows)
    for s in range(0, a.rows, 262144):
        m = min(262144, a.rows - s)
        rows = synth_rows_device(0, SEED_TABLE, nc, s, m)
        ids = np.zeros((m, 2), np.uint64)
        ids[:, 1] = np.arange(s, s + m, dtype=np.uint64)
        ix.upsert(ids, rows)
    torch.cuda.synchronize()
    Q, _ = Synth(nc).queries(16, a.rows)
    k = 12
    qpin = torch.from_numpy(Q).pin_memory()
    qdev = torch.from_numpy(Q).cuda()
    o_ids = np.zeros((1, k, 2), np.uint64); o_d = np.zeros((1, k)); o_c = np.zeros(1, np.int32)
    d_ids = torch.zeros((1, k, 2), dtype=torch.int64, device="cuda"); d_d = torch.zeros((1, k), dtype=torch.float64, device="cuda")
    d_c = torch.zeros(1, dtype=torch.int32, device="cuda")
    h = ix._h

    def loop(kind, timing):
        ix.set_option(ORX_OPT_SCAN_TIMING, timing)
        lat = np.empty(a.iters)
        s0 = ix.stats()
        for i in range(a.iters + 20):
            j = i % 16
            t = time.perf_counter()
            if kind == "host":
                rc = lib.orx_search(h, C.c_void_p(qpin[j].data_ptr()), 1, 1024, k, C.c_void_p(o_ids.ctypes.data),
                                    C.c_void_p(o_d.ctypes.data), C.c_void_p(o_c.ctypes.data))
            else:
                rc = lib.orx_search(h, C.c_void_p(qdev[j].data_ptr()), 1, 1024, k, C.c_void_p(d_ids.data_ptr()),
                                    C.c_void_p(d_d.data_ptr()), C.c_void_p(d_c.data_ptr()))
            if i >= 20:
                lat[i - 20] = time.perf_counter() - t
            check(rc)
        s1 = ix.stats()
        n = s1["scan_launches"] - s0["scan_launches"]
        scan_ms = (s1["scan_ms_total"] - s0["scan_ms_total"]) / n if n else None
        return {"variant": f"{kind} buffers, scan timing {'on' if timing else 'off'}", "p50_us": round(float(np.median(lat)) * 1e6, 2),
                "p10_us": round(float(np.percentile(lat, 10)) * 1e6, 2), "p99_us": round(float(np.percentile(lat, 99)) * 1e6, 2),
                "scan_kernel_us": None if scan_ms is None else round(scan_ms * 1e3, 2),
                "above_scan_us": None if scan_ms is None else round(float(np.median(lat)) * 1e6 - scan_ms * 1e3, 2)}

    out = {"rows": a.rows, "dtype": a.dtype, "devices": devs, "results": [loop("host", 1), loop("device", 1), loop("host", 0), loop("device", 0)]}
    print(json.dumps(out))
    ix.close()


if __name__ == "__main__":
    main()
