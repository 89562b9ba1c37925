"""GPU: the Ward linkage on a thread-block cluster (csrc/hb_group.cuh, ward_cluster_kernel) against the
single-block kernel — merges, their order and heights must be bit-identical — and against scipy."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _linkage(D, mode):
    import torch
    from honeybadger_b200 import _lib, hmm
    lib = _lib.load()
    N = D.shape[0]
    d = torch.from_numpy(np.ascontiguousarray(D)).cuda()
    ma = torch.empty(N - 1, dtype=torch.int32, device="cuda")
    mb = torch.empty(N - 1, dtype=torch.int32, device="cuda")
    hh = torch.empty(N - 1, dtype=torch.float64, device="cuda")
    hmm.set_ward_mode(mode)
    try:
        _lib.check(lib.hb_ward_linkage_dev(d.data_ptr(), N, ma.data_ptr(), mb.data_ptr(), hh.data_ptr(), None))
        torch.cuda.synchronize()
    finally:
        hmm.set_ward_mode(0)
    return ma.cpu().numpy(), mb.cpu().numpy(), hh.cpu().numpy()


def _sqdist(x):
    g = x @ x.T
    n = np.diag(g)
    d = n[:, None] + n[None, :] - 2 * g
    d = np.maximum((d + d.T) / 2, 0.0)
    np.fill_diagonal(d, 0.0)
    return d


@pytest.mark.parametrize("N", [2, 3, 33, 64, 511, 512, 513, 1000, 2049, 4100])
def test_cluster_linkage_is_bit_identical_to_the_single_block_kernel(N):
    rng = np.random.default_rng(N)
    x = rng.normal(size=(N, 12))
    x[: N // 3] += 3.0
    D = _sqdist(x)
    a1, b1, h1 = _linkage(D, 1)
    a2, b2, h2 = _linkage(D, 2)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert np.array_equal(h1.view(np.int64), h2.view(np.int64))


def test_cluster_linkage_with_ties():
    """Integer coordinates: many equal distances, the tie rules (previous chain element, then lowest index)
    decide — the two kernels must still agree merge for merge."""
    rng = np.random.default_rng(5)
    for N in (130, 777, 1500):
        x = rng.integers(0, 3, size=(N, 4)).astype(np.float64)
        D = _sqdist(x)
        a1, b1, h1 = _linkage(D, 1)
        a2, b2, h2 = _linkage(D, 2)
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
        assert np.array_equal(h1.view(np.int64), h2.view(np.int64))


def test_cluster_linkage_matches_scipy_ward_heights():
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import squareform
    rng = np.random.default_rng(9)
    N = 1200
    x = rng.normal(size=(N, 20))
    x[:400] += 2.0
    D = _sqdist(x)
    _, _, h = _linkage(D, 2)
    Z = linkage(squareform(np.sqrt(D), checks=False), method="ward")
    np.testing.assert_allclose(np.sort(h), Z[:, 2], rtol=1e-9, atol=1e-9)


def test_default_mode_on_both_sides_of_the_switch_over(monkeypatch):
    """mode 0 (one block below 256 cells, the cluster from there) must give the single block's answer."""
    rng = np.random.default_rng(11)
    for N in (255, 256, 1023, 1024):
        D = _sqdist(rng.normal(size=(N, 8)))
        a0, b0, h0 = _linkage(D, 0)
        a1, b1, h1 = _linkage(D, 1)
        assert np.array_equal(a0, a1) and np.array_equal(b0, b1) and np.array_equal(h0, h1)


def _dist(S, mode):
    import torch
    from honeybadger_b200 import _lib, hmm
    lib = _lib.load()
    G, N = S.shape
    s = hmm.alloc_rows(G, N, torch.float64, "cuda")
    s.copy_(torch.from_numpy(S).cuda())
    D = torch.empty((N, N), dtype=torch.float64, device="cuda")
    hmm.set_dist_mode(mode)
    try:
        _lib.check(lib.hb_dist_sq_dev(s.data_ptr(), s.stride(0), G, N, D.data_ptr(), None))
        torch.cuda.synchronize()
    finally:
        hmm.set_dist_mode(0)
    return D.cpu().numpy()


@pytest.mark.parametrize("G,N", [(1, 2), (15, 3), (16, 64), (17, 65), (100, 127), (333, 130), (1000, 257), (48, 1)])
def test_pipelined_dist_is_bit_identical_to_the_plain_kernel(G, N):
    rng = np.random.default_rng(G * 1000 + N)
    S = rng.normal(size=(G, N))
    a, b = _dist(S, 0), _dist(S, 1)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))
    ref = ((S[:, :, None] - S[:, None, :]) ** 2).sum(axis=0)
    np.testing.assert_allclose(a, ref, rtol=1e-12, atol=1e-12)


def test_pipelined_dist_with_missing_values():
    """R's NA rule (pairs with a non-finite member dropped, sum scaled by G / pairs used, no pair -> 0): NaN
    scattered, whole NaN columns, NaN confined to some slabs and tiles, +-Inf."""
    rng = np.random.default_rng(3)
    G, N = 200, 150
    S = rng.normal(size=(G, N))
    S[rng.random((G, N)) < 0.02] = np.nan
    S[:, 7] = np.nan
    S[32:48, 64:128] = np.nan
    S[5, 3] = np.inf
    S[6, 4] = -np.inf
    a, b = _dist(S, 0), _dist(S, 1)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))
    d = S[:, :, None] - S[:, None, :]
    ok = np.isfinite(d)
    cnt = ok.sum(axis=0)
    ss = np.where(ok, d * d, 0.0).sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = np.where(cnt > 0, ss / (cnt / G), 0.0)
    np.fill_diagonal(ref, 0.0)
    np.testing.assert_allclose(a, ref, rtol=1e-12, atol=1e-12)


def test_cluster_linkage_and_dist_are_reproducible_run_to_run():
    """Same input, repeated: the cluster kernel's exchange through distributed shared memory and the cp.async ring of
    the distance kernel must not depend on timing (a race would show up as run-to-run differences)."""
    rng = np.random.default_rng(17)
    N, G = 2500, 600
    S = rng.normal(size=(G, N))
    S[:, : N // 2] += 0.7
    S[rng.random((G, N)) < 0.05] = np.nan
    d0 = _dist(S, 0)
    first = _linkage(d0, 2)
    for _ in range(4):
        d = _dist(S, 0)
        assert np.array_equal(d.view(np.int64), d0.view(np.int64))
        again = _linkage(d, 2)
        assert all(np.array_equal(a.view(np.int64) if a.dtype == np.float64 else a,
                                  b.view(np.int64) if b.dtype == np.float64 else b) for a, b in zip(first, again))
    plain = _linkage(d0, 1)
    assert np.array_equal(first[0], plain[0]) and np.array_equal(first[1], plain[1])
    assert np.array_equal(first[2].view(np.int64), plain[2].view(np.int64))


@pytest.mark.parametrize("N,r", [(1100, 1), (1100, 2), (4000, 4), (6000, 4), (8000, 4), (9000, 4)])
def test_every_instantiation_of_the_cluster_kernel(N, r, monkeypatch):
    """The cluster kernel is compiled for 1, 2, 3, 4, 6, 8 and 16 columns per thread; HB_WARD_R forces small
    clusters so that each instantiation runs (5, 3, 4, 6, 8, 9 -> 16 columns here) and must reproduce the single
    block merge for merge."""
    rng = np.random.default_rng(N + r)
    x = rng.normal(size=(N, 6))
    x[: N // 3] += 1.5
    D = _sqdist(x)
    monkeypatch.setenv("HB_WARD_R", str(r))
    a2, b2, h2 = _linkage(D, 2)
    monkeypatch.delenv("HB_WARD_R")
    a1, b1, h1 = _linkage(D, 1)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert np.array_equal(h1.view(np.int64), h2.view(np.int64))
