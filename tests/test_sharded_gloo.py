"""world_size-2 gloo test of the row-shard + exchange + merge host logic (no GPU):
the local table and the merge are injected oracle-backed fakes, the partitioning, packing,
all_gather and result layout are the product's own (outline_rag_b200/sharded.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleLocalIndex:
    """CPU stand-in with Index's interface, for the gloo test only."""

    def __init__(self):
        self.ids = np.zeros((0, 2), np.uint64)
        self.X = np.zeros((0, 1024), np.float32)

    def __len__(self):
        return self.ids.shape[0]

    def upsert(self, ids, vecs):
        self.delete(ids)
        self.ids = np.concatenate([self.ids, np.asarray(ids, np.uint64).reshape(-1, 2)])
        self.X = np.concatenate([self.X, np.asarray(vecs, np.float32)])

    def delete(self, ids):
        kill = {tuple(map(int, r)) for r in np.asarray(ids, np.uint64).reshape(-1, 2)}
        keep = np.array([tuple(map(int, r)) not in kill for r in self.ids], bool)
        removed = int((~keep).sum())
        self.ids, self.X = self.ids[keep], self.X[keep]
        return removed

    def search(self, queries, k):
        from oracle import cosine_topk as O
        nq = queries.shape[0]
        ids = np.zeros((nq, k, 2), np.uint64)
        d = np.full((nq, k), np.nan)
        cnt = np.zeros(nq, np.int32)
        for i in range(nq):
            a, b = O.topk_exact(self.X, self.ids, queries[i], k)
            ids[i, :len(b)], d[i, :len(b)], cnt[i] = a, b, len(b)
        return ids, d, cnt


    def search_filtered(self, queries, k, allow):
        ok = {tuple(map(int, r)) for r in np.asarray(allow, np.uint64).reshape(-1, 2)}
        keep = np.array([tuple(map(int, r)) in ok for r in self.ids], bool)
        sub = OracleLocalIndex()
        sub.ids, sub.X = self.ids[keep], self.X[keep]
        self.last_allow_size = len(ok)
        return sub.search(np.asarray(queries, np.float32), k)


def oracle_merge(g_ids, g_dist, g_cnt, k):
    from oracle import cosine_topk as O
    n_lists, nq = g_dist.shape[:2]
    ids = np.zeros((nq, k, 2), np.uint64)
    d = np.full((nq, k), np.nan)
    cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        parts = [(g_ids[l, i, :g_cnt[l, i]], g_dist[l, i, :g_cnt[l, i]]) for l in range(n_lists)]
        a, b = O.merge_shards(parts, k)
        ids[i, :len(b)], d[i, :len(b)], cnt[i] = a, b, len(b)
    return ids, d, cnt


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cosine_topk as O
        from outline_rag_b200.sharded import ShardedIndex, shard_of
        from orx_testkit.synth import Synth
        syn = Synth(64)
        n, k = 1500, 12
        X = syn.table(n)
        Q, _ = syn.queries(5, n)
        ids = O.ids_arange(0, n)
        sh = ShardedIndex(local_index=OracleLocalIndex(), merge_fn=oracle_merge)
        kept = sh.upsert(ids, X)
        assert kept == int((shard_of(ids, world) == rank).sum())
        assert sh.global_size() == n
        got = sh.search(Q, k)
        # delete a doc's worth of rows everywhere, then search again
        gone = ids[100:140]
        sh.delete(gone)
        assert sh.global_size() == n - 40
        got2 = sh.search(Q, k)
        # the WHERE-clause search: every rank is handed only the allowed ids it owns
        allow = ids[200:260]
        got3 = sh.search_filtered(Q, k, allow)
        assert sh.local.last_allow_size == int((shard_of(allow, world) == rank).sum())
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), ids=got[0], d=got[1], c=got[2], ids2=got2[0],
                 d2=got2[1], c2=got2[2], ids3=got3[0], d3=got3[1], c3=got3[2], kept=kept)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharded_search_equals_single_table(tmp_path):
    from oracle import cosine_topk as O
    from orx_testkit.synth import Synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    syn = Synth(64)
    n, k = 1500, 12
    X = syn.table(n)
    Q, _ = syn.queries(5, n)
    ids = O.ids_arange(0, n)
    r = [np.load(tmp_path / f"r{i}.npz") for i in range(world)]
    assert int(r[0]["kept"]) + int(r[1]["kept"]) == n
    assert abs(int(r[0]["kept"]) - n / 2) < 0.1 * n          # balanced partition
    keep = np.ones(n, bool)
    keep[100:140] = False
    for qi in range(Q.shape[0]):
        w_ids, w_d = O.topk_exact(X, ids, Q[qi], k)
        w2_ids, w2_d = O.topk_exact(X[keep], ids[keep], Q[qi], k)
        for rr in r:                                         # every rank holds the global answer
            assert np.array_equal(rr["ids"][qi], w_ids) and np.array_equal(rr["d"][qi], w_d)
            assert rr["c"][qi] == k
            assert np.array_equal(rr["ids2"][qi], w2_ids) and np.array_equal(rr["d2"][qi], w2_d)
        w3_ids, w3_d = O.topk_exact(X[200:260], ids[200:260], Q[qi], k, exhaustive=True)
        for rr in r:
            assert rr["c3"][qi] == k and np.array_equal(rr["ids3"][qi], w3_ids) and np.array_equal(rr["d3"][qi], w3_d)


def test_pack_unpack_roundtrip():
    from outline_rag_b200.sharded import pack_results, unpack_results
    rng = np.random.default_rng(0)
    ids = torch.from_numpy(rng.integers(-2**62, 2**62, size=(3, 12, 2)))
    d = torch.from_numpy(rng.standard_normal((3, 12)))
    d[1, 5:] = float("nan")
    c = torch.tensor([12, 5, 12], dtype=torch.int32)
    block = pack_results(ids, d, c)
    assert block.shape == (3, 37)
    g = torch.stack([block, block])
    i2, d2, c2 = unpack_results(g, 12)
    assert torch.equal(i2[1], ids) and torch.equal(c2[0], c)
    assert torch.equal(d2[0].view(torch.int64), d.view(torch.int64))
