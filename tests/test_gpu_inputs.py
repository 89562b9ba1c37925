"""Input construction on the device against the oracle: bit-exact, through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a, dtype):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device="cuda")


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("shape", [(1, 1), (33, 65), (257, 91), (700, 1300)])
def test_row_means_and_column_stats_bit_exact(oracle, shape):
    import torch
    from honeybadger_b200 import _lib
    lib = _lib.load()
    G, N = shape
    rng = np.random.default_rng(G * 7 + N)
    m = rng.gamma(2.0, 1.5, (G, N)) * (rng.random((G, N)) > 0.3)
    ld = N + 5
    buf = torch.zeros((G, ld), dtype=torch.float64, device="cuda")
    buf[:, :N] = _dev(m, np.float64)
    out = torch.zeros(G, dtype=torch.float64, device="cuda")
    _lib.check(lib.hb_row_means_dev(buf.data_ptr(), ld, G, N, out.data_ptr(), _stream()))
    assert np.array_equal(out.cpu().numpy(), oracle.row_means(m))
    keep = (rng.random(G) > 0.25).astype(np.uint8)
    keep[0] = 1
    cen = torch.zeros(N, dtype=torch.float64, device="cuda")
    scl = torch.zeros(N, dtype=torch.float64, device="cuda")
    kd = _dev(keep, np.uint8)
    _lib.check(lib.hb_col_scale_stats_dev(buf.data_ptr(), ld, G, N, kd.data_ptr(), cen.data_ptr(),
                                          scl.data_ptr(), _stream()))
    sub = m[keep.astype(bool)]
    want_c = oracle.row_means(sub.T)  # colMeans = rowMeans of the transpose (same accumulation order)
    assert np.array_equal(cen.cpu().numpy(), want_c)
    if sub.shape[0] > 1 and (sub.std(0) > 0).all():
        scaled = oracle.scale(sub)
        rows = _dev(np.nonzero(keep)[0], np.int32)
        o = torch.zeros((sub.shape[0], N), dtype=torch.float64, device="cuda")
        _lib.check(lib.hb_gexp_apply_dev(buf.data_ptr(), ld, N, rows.data_ptr(), sub.shape[0], cen.data_ptr(),
                                         scl.data_ptr(), None, o.data_ptr(), N, _stream()))
        assert np.array_equal(o.cpu().numpy(), scaled)


def test_mean_nonzero_fast_and_exact(oracle):
    from honeybadger_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(8)
    m = rng.gamma(2.0, 1.5, (311, 77)) * (rng.random((311, 77)) > 0.6)
    d = _dev(m, np.float64)
    want = oracle.mean_nonzero(m)
    mean, margin = C.c_double(0), C.c_double(-1)
    _lib.check(lib.hb_mean_nonzero_dev(d.data_ptr(), 77, 311, 77, 1, C.byref(mean), C.byref(margin), _stream()))
    assert mean.value == want and margin.value == 0.0
    _lib.check(lib.hb_mean_nonzero_dev(d.data_ptr(), 77, 311, 77, 0, C.byref(mean), C.byref(margin), _stream()))
    assert abs(mean.value - want) <= margin.value and 0 < margin.value < 1e-12
    z = _dev(np.zeros((4, 4)), np.float64)
    _lib.check(lib.hb_mean_nonzero_dev(z.data_ptr(), 4, 4, 4, 0, C.byref(mean), None, _stream()))
    assert np.isnan(mean.value)  # mean(numeric(0))


@pytest.mark.parametrize("filt,scale", [(False, False), (False, True), (True, False), (True, True)])
def test_set_gexp_mats_dev_and_host_match_oracle(oracle, golden_inputs, filt, scale):
    from honeybadger_b200 import _lib
    from honeybadger_b200.honeybadger import set_gexp_mats_dev
    gexp, ref = golden_inputs["gexp"], golden_inputs["ref"][:, None]
    rng = np.random.default_rng(1)
    ref3 = ref + rng.normal(0, 0.3, (ref.shape[0], 3))  # a 3-column reference exercises scale(ref) + rowMeans
    for rf in (ref, ref3):
        want, keep = oracle.set_gexp_mats(gexp, rf, filter=filt, scale=scale)
        norm, k, info = set_gexp_mats_dev(gexp, rf, filter=filt, scale=scale)
        assert np.array_equal(k, keep) and info == 0
        assert np.array_equal(norm.cpu().numpy(), want)
        # explicit thresholds (the vignette-commented values of R/HoneyBADGER.R:122)
        if filt:
            want2, keep2 = oracle.set_gexp_mats(gexp, rf, True, 4.5, 6.0, 8.0, scale)
            norm2, k2, _ = set_gexp_mats_dev(gexp, rf, True, 4.5, 6.0, 8.0, scale)
            assert np.array_equal(k2, keep2) and np.array_equal(norm2.cpu().numpy(), want2)
    # host form, matrices as R holds them (column-major)
    lib = _lib.load()
    G, N = gexp.shape
    sc_f, ref_f = np.asfortranarray(gexp), np.asfortranarray(ref3)
    out = np.zeros(G * N)
    kr = np.zeros(G, np.int32)
    gk, info = C.c_int64(0), C.c_int32(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.hb_set_gexp_mats_host(vp(sc_f), G, N, vp(ref_f), 3, int(filt), 0.0, None, None, int(scale),
                                         vp(out), vp(kr), C.byref(gk), C.byref(info)))
    want, keep = oracle.set_gexp_mats(gexp, ref3, filter=filt, scale=scale)
    assert gk.value == keep.size and np.array_equal(kr[:gk.value], keep)
    assert np.array_equal(out[:gk.value * N].reshape(N, gk.value).T, want)


def test_set_gexp_mats_threshold_tie_replays_r_mean(oracle):
    """A row mean equal to the default threshold: only R's serially rounded mean(x[x != 0]) decides the
    comparison, so the library must fall back to the exact replay (info bit 0) and agree with the oracle."""
    from honeybadger_b200.honeybadger import set_gexp_mats_dev
    rng = np.random.default_rng(12)
    G, N = 64, 48
    sc = np.zeros((G, N))
    sc[:32] = 2.0                      # every non-zero entry is 2 -> the threshold is exactly 2 ...
    sc[32:48] = np.where(rng.random((16, N)) > 0.5, 2.0, 0.0)
    ref = np.full((G, 1), -1.0)        # ... rows 0..31 have row mean exactly 2: not > 2, and ref never passes
    ref[60:] = 0.0
    want, keep = oracle.set_gexp_mats(sc, ref, filter=True, scale=False)
    norm, k, info = set_gexp_mats_dev(sc, ref, filter=True, scale=False)
    assert info & 1
    assert np.array_equal(k, keep) and np.array_equal(norm.cpu().numpy(), want)


def test_set_allele_mats_known_answers(oracle, golden_inputs, golden_allele):
    from honeybadger_b200 import _lib
    from honeybadger_b200.honeybadger import set_allele_mats_dev
    r, cov = golden_inputs["r"].astype(np.int32), golden_inputs["cov"].astype(np.int32)
    for thr, (printed, final) in zip((0.05, 0.1), golden_inputs["known_counts"]):
        out = set_allele_mats_dev(r, cov, het_deviance_threshold=thr)
        want = oracle.set_allele_mats(r, cov, het_deviance_threshold=thr)
        assert out["n_het"] == printed and out["keep"].size == final  # the reference's own printouts
        for key in ("l", "n_bulk", "l_maf", "keep"):
            assert np.array_equal(out[key], want[key])
        assert np.array_equal(out["r_maf"].cpu().numpy(), want["r_maf"])
        assert np.array_equal(out["n_sc"].cpu().numpy(), want["n_sc"])
    assert np.array_equal(out["r_maf"].cpu().numpy(), golden_allele["r_maf"].astype(np.int32))
    # no filter, user-supplied bulk, NA rows; host form in R layout
    rng = np.random.default_rng(2)
    n = rng.integers(0, 4, (90, 37)).astype(np.int32)
    n[11] = 0
    rr = rng.binomial(n, 0.6).astype(np.int32)
    li, ni = rng.integers(0, 30, 90).astype(float), rng.integers(1, 60, 90).astype(float)
    for kw in (dict(filter=False), dict(filter=True, l_init=li, n_bulk_init=ni, min_cell=2)):
        out = set_allele_mats_dev(rr, n, **kw)
        want = oracle.set_allele_mats(rr, n, **kw)
        for key in ("l", "n_bulk", "l_maf", "keep"):
            assert np.array_equal(out[key], want[key])
        assert np.array_equal(out["r_maf"].cpu().numpy(), want["r_maf"])
    lib = _lib.load()
    G, N = rr.shape
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rf, nf = np.asfortranarray(rr), np.asfortranarray(n)
    o_r, o_n = np.zeros(G * N, np.int32), np.zeros(G * N, np.int32)
    l, nb, lm, kr = np.zeros(G), np.zeros(G), np.zeros(G), np.zeros(G, np.int32)
    gk, nh = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.hb_set_allele_mats_host(vp(rf), vp(nf), G, N, None, None, 1, 0.05, 3, vp(o_r), vp(o_n), vp(l),
                                           vp(nb), vp(lm), vp(kr), C.byref(gk), C.byref(nh)))
    want = oracle.set_allele_mats(rr, n)
    k = gk.value
    assert k == want["keep"].size and np.array_equal(kr[:k], want["keep"])
    assert np.array_equal(o_r[:k * N].reshape(N, k).T, want["r_maf"])
    assert np.array_equal(o_n[:k * N].reshape(N, k).T, want["n_sc"])
    assert np.array_equal(lm[:k], want["l_maf"])


def test_refclass_set_mats_stay_resident(oracle, golden_inputs):
    """HoneyBADGER$setGexpMats / setAlleleMats mirrors: same result as the oracle, matrices left on the GPU."""
    import pandas as pd
    from honeybadger_b200.honeybadger import GRanges, HoneyBADGER
    gexp, ref = golden_inputs["gexp"], golden_inputs["ref"]
    G, N = gexp.shape
    gn = [f"g{i}" for i in range(G)]
    cells = [f"c{i}" for i in range(N)]
    pos_ok = np.ones(G, bool)
    pos_ok[5::17] = False  # genes without a position are dropped after the normalisation
    names = [g for g, ok in zip(gn, pos_ok) if ok]
    genes = GRanges(seqnames=["chr%d" % (1 + i % 22) for i in range(len(names))],
                    start=np.arange(len(names)) * 10, end=np.arange(len(names)) * 10 + 5, names=names)
    hb = HoneyBADGER("t")
    hb.setGexpMats(pd.DataFrame(gexp, index=gn, columns=cells), pd.DataFrame(ref[:, None], index=gn), genes,
                   filter=True, scale=True, verbose=False)
    want, keep = oracle.set_gexp_mats(gexp, ref[:, None], filter=True, scale=True)
    sel = pos_ok[keep]
    assert list(hb.gexp_norm.index) == [gn[i] for i in keep[sel]]
    assert np.array_equal(hb.gexp_norm.values, want[sel])
    assert hb._x_dev is not None and np.array_equal(hb._x_dev[0].cpu().numpy(), want[sel])
    r, cov = golden_inputs["r"].astype(np.int32), golden_inputs["cov"].astype(np.int32)
    sn = [f"{1 + i % 22}:{1000 + i}" for i in range(r.shape[0])]
    hb.setAlleleMats(pd.DataFrame(r, index=sn), pd.DataFrame(cov, index=sn), het_deviance_threshold=0.1,
                     verbose=False)
    wa = oracle.set_allele_mats(r, cov, het_deviance_threshold=0.1)
    assert np.array_equal(hb.r_maf.values, wa["r_maf"]) and np.array_equal(hb.n_sc.values, wa["n_sc"])
    assert np.array_equal(hb.l_maf, wa["l_maf"]) and hb._r_dev is not None


def test_set_gexp_mats_edge_shapes_and_values(oracle):
    """Ragged and degenerate inputs: one gene, one cell, constant columns (scale() divides 0 by 0 -> NaN in R
    too), negative and large-magnitude values, every row filtered out."""
    from honeybadger_b200.honeybadger import set_gexp_mats_dev
    rng = np.random.default_rng(77)
    cases = []
    for G, N, R in ((1, 1, 1), (1, 9, 2), (5, 1, 1), (33, 2, 3), (130, 67, 1), (64, 300, 5)):
        sc = rng.normal(0, 3, (G, N)) * (rng.random((G, N)) > 0.2)
        ref = rng.normal(0, 2, (G, R))
        cases.append((sc, ref))
    big = rng.normal(0, 1, (40, 17)) * 1e12
    cases.append((big, rng.normal(0, 1, (40, 2)) * 1e-9))
    const = rng.normal(0, 1, (20, 6))
    const[:, 2] = 3.25  # a constant cell: centred to 0, scale 0
    cases.append((const, rng.normal(0, 1, (20, 1))))
    for sc, ref in cases:
        for filt, scale in ((False, False), (False, True), (True, True)):
            want, keep = oracle.set_gexp_mats(sc, ref, filter=filt, scale=scale)
            with np.errstate(all="ignore"):
                norm, k, _ = set_gexp_mats_dev(sc, ref, filter=filt, scale=scale)
            assert np.array_equal(k, keep)
            assert np.array_equal(norm.cpu().numpy(), want, equal_nan=True)
    # nothing passes the filter
    sc = -np.abs(rng.normal(0, 1, (12, 5))) - 1.0
    norm, k, _ = set_gexp_mats_dev(sc, sc[:, :1], filter=True, minMeanBoth=0.0, minMeanTest=0.0, minMeanRef=0.0,
                                   scale=False)
    assert k.size == 0 and norm.shape[0] == 0
