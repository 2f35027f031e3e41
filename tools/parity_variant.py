#!/usr/bin/env python
"""Oracle parity of an experimental liborx build (make -C outline_rag_b200/csrc variant NAME=x DEFS=...), batch paths:
    python tools/parity_variant.py --lib outline_rag_b200/liborx_x.so
Checks ids AND distance bits against oracle.topk_exact for batches that select each tcgen05 kernel (1 CTA, CTA pairs,
clusters of 4 when the build enables them), both table dtypes, ragged row counts."""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--rows", type=int, default=40_003)
    a = ap.parse_args()
    if a.lib:
        sys.modules["orx_lib_override"] = types.SimpleNamespace(LIB_PATH=os.path.abspath(a.lib))
    import outline_rag_b200 as orx
    from oracle import cosine_topk as O
    from oracle.synth_host import FastSynth
    from tests._helpers import stored_bf16_rows
    syn = FastSynth(1024)
    n, k = a.rows, 12
    X = syn.table(n).copy()
    X[17] = 0.0
    X[99] *= np.float32(1e-30)
    Q, _ = syn.queries(1024, n)
    ids = O.ids_arange(0, n)
    bad = 0
    for dtype in ("fp32", "bf16"):
        rows = X if dtype == "fp32" else stored_bf16_rows(X)
        with orx.Index(dtype, device=0) as ix:
            ix.upsert(ids, X)
            for nq in (5, 200, 512, 1024, 700):
                g_ids, g_d, g_c = ix.search(Q[:nq], k)
                st = ix.stats()
                wrong = 0
                for i in range(0, nq, max(1, nq // 64)):
                    w_ids, w_d = O.topk_exact(rows, ids, Q[i], k)
                    if not (np.array_equal(g_ids[i], w_ids) and np.array_equal(g_d[i].view(np.uint64), w_d.view(np.uint64))):
                        wrong += 1
                bad += wrong
                print(f"{dtype} nq={nq}: path {st['last_path']} wrong {wrong} fallbacks gemv {st['fallback_gemv']} exhaustive {st['fallback_exhaustive']}", flush=True)
    print("PARITY", "OK" if bad == 0 else f"FAILED ({bad})")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
