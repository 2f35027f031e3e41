"""The row-sharded search over NVLink peer memory (orx_search_sharded).

* world = 1 on any B200: the publish / merge-wait chain degenerates to one rank and must equal
  orx_search (runs in the driver's 1-GPU `-m gpu` pass).
* world = 2 with one process per GPU (NCCL only for the one-off handle exchange): needs >= 2 GPUs,
  skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`.
"""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import cosine_topk as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = 12


def test_single_rank_exchange_equals_plain_search(small_table):
    import torch
    import outline_rag_b200 as orx
    X, Q, _ = small_table
    ids = O.ids_arange(0, 5000)
    with orx.Index("fp32") as ix:
        ix.upsert(ids, X[:5000])
        ix.shard_connect([ix.shard_export(1, 0)])
        for nq in (1, 5, 40):
            a = ix.search(Q[:nq], K)
            b = ix.search_sharded(Q[:nq], K)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))
            assert np.array_equal(a[2], b[2])
            d = ix.search_sharded(torch.from_numpy(Q[:nq]).cuda(), K)
            assert np.array_equal(d[0].cpu().numpy().view(np.uint64), a[0])
        # flagged queries (exact ties wider than the candidate list) take the redo round
        dup = np.tile(X[7], (300, 1))
        ix.upsert(O.ids_arange(10000, 10300), dup)
        a = ix.search(X[7], K)
        b = ix.search_sharded(X[7], K)
        assert np.array_equal(a[0], b[0]) and O.ids_to_ints(b[0][0])[0] == 7
        bad = Q[:2].copy()
        bad[0, 3] = np.inf
        with pytest.raises(orx.OrxValueError, match="NaN or infinite"):
            ix.search_sharded(bad, K)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from outline_rag_b200.sharded import ShardedIndex
        from orx_testkit.synth import Synth
        syn = Synth(1024)
        n = 30000
        X = syn.table(n)
        Q, _ = syn.queries(70, n)
        ids = O.ids_arange(0, n)
        res = {}
        for mode in ("p2p", "nccl"):
            os.environ["ORX_SHARD_EXCHANGE"] = mode
            sh = ShardedIndex("fp32", n, rank)
            assert sh.exchange == mode, (sh.exchange, getattr(sh, "_p2p_error", None))
            sh.upsert(ids, X)
            assert sh.global_size() == n
            qd = torch.from_numpy(Q).cuda()
            for nq in (1, 70):
                got = sh.search(qd[:nq], K)
                res[f"{mode}_ids_{nq}"] = got[0].cpu().numpy().view(np.uint64).copy()
                res[f"{mode}_d_{nq}"] = got[1].cpu().numpy().copy()
            if mode == "nccl":
                # the WHERE-clause search across shards (each rank resolves the ids it owns; all_gather + merge)
                f = sh.search_filtered(Q[:3], K, ids[5000:15000])
                res["filtered_ids"] = f[0]
                res["filtered_d"] = f[1]
            if mode == "p2p":
                h = sh.search(Q[:3], K)                       # host buffers through the C-ABI
                res["p2p_host_ids"] = h[0]
                # two collective searches in flight (orx_search_sharded_submit / orx_search_wait), incl. a query whose
                # shard-local candidates are a flood of exact ties (second, exact round while another search is in flight)
                sh.upsert(O.ids_arange(900_000, 900_300), np.tile(X[7], (300, 1)))
                qs = torch.from_numpy(np.concatenate([Q[:5], X[7:8], Q[5:8]])).cuda()
                piped, prev = [], None
                for i in range(qs.shape[0]):
                    t, out = sh.search_submit(qs[i:i + 1], K)
                    if prev is not None:
                        sh.search_wait(prev[0])
                        piped.append(prev[1][0].cpu().numpy().view(np.uint64).copy())
                    prev = (t, out)
                sh.search_wait(prev[0])
                piped.append(prev[1][0].cpu().numpy().view(np.uint64).copy())
                res["p2p_piped_ids"] = np.concatenate(piped)
                sh.delete(O.ids_arange(900_000, 900_300))
                sh.delete(ids[100:160])
                got = sh.search(qd[:4], K)
                res["p2p_after_delete_ids"] = got[0].cpu().numpy().view(np.uint64).copy()
            sh.local.close()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


def test_two_gpus_p2p_and_nccl_equal_the_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from orx_testkit.synth import Synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    syn = Synth(1024)
    n = 30000
    X = syn.table(n)
    Q, _ = syn.queries(70, n)
    ids = O.ids_arange(0, n)
    r = [np.load(tmp_path / f"r{i}.npz") for i in range(2)]
    keep = np.ones(n, bool)
    keep[100:160] = False
    dupX = np.concatenate([X, np.tile(X[7], (300, 1))])
    dup_ids = np.concatenate([ids, O.ids_arange(900_000, 900_300)])
    w7, _ = O.topk_exact(dupX, dup_ids, X[7], K)
    for rr in r:
        assert np.array_equal(rr["p2p_piped_ids"][5], w7), "the query that needed the exact second round"
    for qi in range(70):
        w_ids, w_d = O.topk_exact(X, ids, Q[qi], K)
        for rr in r:
            for mode in ("p2p", "nccl"):
                assert np.array_equal(rr[f"{mode}_ids_70"][qi], w_ids)
                assert np.array_equal(rr[f"{mode}_d_70"][qi].view(np.uint64), w_d.view(np.uint64))
            if qi == 0:
                assert np.array_equal(rr["p2p_ids_1"][0], w_ids) and np.array_equal(rr["nccl_ids_1"][0], w_ids)
            if qi < 3:
                assert np.array_equal(rr["p2p_host_ids"][qi], w_ids)
            if qi < 8:
                wp, _ = O.topk_exact(dupX, dup_ids, Q[qi], K)       # (the tie flood is in the table during these searches)
                assert np.array_equal(rr["p2p_piped_ids"][qi if qi < 5 else qi + 1], wp), ("two in flight", qi)
            if qi < 4:
                w2, _ = O.topk_exact(X[keep], ids[keep], Q[qi], K)
                assert np.array_equal(rr["p2p_after_delete_ids"][qi], w2)
            if qi < 3:
                w3, w3d = O.topk_exact(X[5000:15000], ids[5000:15000], Q[qi], K)
                assert np.array_equal(rr["filtered_ids"][qi], w3)
                assert np.array_equal(rr["filtered_d"][qi].view(np.uint64), w3d.view(np.uint64))
