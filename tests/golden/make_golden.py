"""Regenerates tests/golden/topk_golden.json.

The reference has no golden vectors for this path, so these are produced by the
oracle itself on the counter-based synthetic table (outline_rag_b200/synth.py): inputs are
reproducible from seeds, only the expected ids + canonical distances (hex floats) are stored.
The GPU parity tests compare the CUDA path against the same file.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cosine_topk as O            # noqa: E402
from orx_testkit.synth import Synth       # noqa: E402
from tests._helpers import stored_bf16_rows          # noqa: E402

N_ROWS, N_QUERIES, K = 20000, 16, 12


def main():
    syn = Synth(1024)
    X = syn.table(N_ROWS)
    Q, anchors = syn.queries(N_QUERIES, N_ROWS)
    ids = O.ids_arange(0, N_ROWS)
    Xb = stored_bf16_rows(X)
    cases = []
    for qi in range(N_QUERIES):
        g_ids, g_d = O.topk_exact(X, ids, Q[qi], K, exhaustive=True)
        b_ids, b_d = O.topk_exact(Xb, ids, Q[qi], K, exhaustive=True)
        cases.append({"anchor": int(anchors[qi]), "ids": O.ids_to_ints(g_ids),
                      "dist_hex": [float(d).hex() for d in g_d],
                      "bf16_ids": O.ids_to_ints(b_ids), "bf16_dist_hex": [float(d).hex() for d in b_d]})
    out = {"n_rows": N_ROWS, "n_centres": 1024, "k": K, "probe_x_123_45": float(X[123, 45]).hex(),
           "cases": cases}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "topk_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
