"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck):
    gpurun -- compute-sanitizer --tool memcheck python tools/sanitize_case.py
Covers upsert, delete (compaction), GEMV scan, tcgen05 scan (both dtypes), the exhaustive fallback,
the single-rank sharded exchange, snapshot export/import, the COPY BINARY decode (odd payload alignment) and
the bitmap-filtered scan; checks results against the oracle."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import outline_rag_b200 as orx                      # noqa: E402
from oracle import cosine_topk as O                 # noqa: E402
from orx_testkit.synth import Synth            # noqa: E402
from tests._helpers import stored_bf16_rows         # noqa: E402

syn = Synth(64)
n = 6000
X = syn.table(n)
Q, _ = syn.queries(40, n)
ids = O.ids_arange(0, n)
keep = np.ones(n, bool)
keep[50:90] = False
for dtype in ("fp32", "bf16"):
    rows = X if dtype == "fp32" else stored_bf16_rows(X)
    with orx.Index(dtype, capacity=1024) as ix:
        ix.upsert(ids, X)
        ix.delete(ids[50:90])
        for nq in (1, 40):
            g = ix.search(Q[:nq], 12)
            for i in range(nq):
                w_ids, w_d = O.topk_exact(rows[keep], ids[keep], Q[i], 12)
                assert np.array_equal(g[0][i], w_ids) and np.array_equal(g[1][i].view(np.uint64), w_d.view(np.uint64))
        ix.upsert(O.ids_arange(10000, 10100), np.tile(X[7], (100, 1)))          # ties -> exhaustive fallback
        g = ix.search(X[7], 12)
        assert ix.stats()["fallback_exhaustive"] >= 1
        ix.shard_connect([ix.shard_export(1, 0)])
        s = ix.search_sharded(Q[:5], 12)
        p = ix.search(Q[:5], 12)
        assert np.array_equal(s[0], p[0])
        with tempfile.TemporaryDirectory() as d:
            ix.save(d)
            jx = orx.Index.load(d)
            assert np.array_equal(jx.search(Q[:3], 12)[0], p[0][:3])
            jx.close()
        # cold-start decode: a 3-byte header extension shifts the payloads to byte alignments 0 and 2 (tests use 1 and 3)
        stream = orx.pgwire.COPY_HEADER[:15] + (3).to_bytes(4, "big") + b"xyz" + orx.pgwire.encode_tuples(ids[:300], X[:300]) \
            + orx.pgwire.COPY_TRAILER
        with orx.Index(dtype) as kx:
            assert kx.load_pgcopy([stream[:100_001], stream[100_001:]]) == (300, 0)
            got, found = kx.fetch(ids[:300])
            assert found.all() and np.array_equal(got, rows[:300])
        # bitmap-filtered scan over 4500 of the live rows (+ the handle form)
        live = np.nonzero(keep)[0]
        sel = live[:4500]
        f = ix.search_filtered(Q[:3], 12, ids[sel])
        with ix.make_filter(ids[sel]) as flt:
            h = ix.search_filtered(Q[:3], 12, flt)
        for i in range(3):
            w_ids, w_d = O.topk_exact(rows[sel], ids[sel], Q[i], 12)
            assert np.array_equal(f[0][i], w_ids) and np.array_equal(h[0][i], w_ids)
            assert np.array_equal(f[1][i].view(np.uint64), w_d.view(np.uint64))
print("sanitize case ok")
