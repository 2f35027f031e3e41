"""Synthetic bge-m3-shaped data: host generator properties, and on the GPU
that csrc/synth.cu produces the SAME BITS (so 10M-row tables can be built in HBM and verified
by regenerating sampled rows on the host)."""
import numpy as np
import pytest

from orx_testkit.synth import Synth, doc_chunk_counts, default_centres


def test_rows_are_unit_norm_and_reproducible(synth100k):
    X = synth100k.table(512)
    n = np.sqrt((X.astype(np.float64) ** 2).sum(1))
    np.testing.assert_allclose(n, 1.0, atol=2e-7)
    again = synth100k.rows(np.array([5, 300, 17], np.uint64))
    assert np.array_equal(again, X[[5, 300, 17]])
    assert np.array_equal(synth100k.table(16, start=100), X[100:116])


def test_statistics_look_like_bge_m3(small_table):
    X, Q, anchors = small_table
    S = X[:2048] @ X[:2048].T
    off = S[~np.eye(2048, dtype=bool)]
    assert 0.15 < off.mean() < 0.35          # anisotropic: random pairs are positively correlated
    sims = (Q @ X.T)
    top = np.sort(sims, axis=1)[:, ::-1]
    assert 0.8 < top[:, 0].mean() < 0.97     # anchor row is the top hit
    assert (sims.argmax(1) == anchors).all()


def test_queries_are_prefix_stable(synth100k):
    a, _ = synth100k.queries(8, 8192)
    b, _ = synth100k.queries(4, 8192)
    assert np.array_equal(a[:4], b)


def test_doc_chunk_counts():
    c = doc_chunk_counts(5000)
    assert c.min() >= 8 and c.max() <= 40 and 17 < c.mean() < 23
    assert default_centres(100_000) == 1024 and default_centres(1_000_000) == 16384


def test_c_host_generator_is_bit_identical(synth100k):
    """oracle/synth_host.c (what the CPU arm of bench.py and the 1M-row tests build their tables with) produces
    the same bits as the NumPy generator: mean, centres, arbitrary rows, contiguous tables, queries."""
    from oracle.synth_host import FastSynth
    fast = FastSynth(synth100k.n_centres)
    assert np.array_equal(fast.mean.view(np.uint32), synth100k.mean.view(np.uint32))
    assert np.array_equal(fast.centres.view(np.uint32), synth100k.centres.view(np.uint32))
    assert np.array_equal(fast.table(2500, start=77).view(np.uint32), synth100k.table(2500, start=77).view(np.uint32))
    idx = np.array([5, 99_999_999, 123, 2**40 + 17], np.uint64)
    assert np.array_equal(fast.rows(idx).view(np.uint32), synth100k.rows(idx).view(np.uint32))
    qa, aa = fast.queries(5, 100_000)
    qb, ab = synth100k.queries(5, 100_000)
    assert np.array_equal(qa.view(np.uint32), qb.view(np.uint32)) and np.array_equal(aa, ab)
    assert FastSynth(16384, threads=3).table(7, start=9_999_990).shape == (7, 1024)


@pytest.mark.gpu
def test_device_generator_is_bit_identical(synth100k):
    import torch
    from orx_testkit.device import synth_rows_device
    dev = synth_rows_device(0, synth100k.seed, synth100k.n_centres, 1000, 4096).cpu().numpy()
    host = synth100k.table(4096, start=1000)
    assert np.array_equal(dev.view(np.uint32), host.view(np.uint32))
    big = Synth(16384)
    dev = synth_rows_device(0, big.seed, big.n_centres, 9_999_000, 64).cpu().numpy()
    assert np.array_equal(dev.view(np.uint32), big.table(64, start=9_999_000).view(np.uint32))
