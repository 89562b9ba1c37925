"""Oracle restatement of the input construction (setGexpMats / setAlleleMats arithmetic)
against independent numpy evaluations and the reference's printed known answers.  CPU only."""
from __future__ import annotations

import numpy as np


def _ld_seq_sum(a, axis):
    """sequential (index-order) accumulation in x87 extended precision: cumsum never re-associates"""
    return np.cumsum(a.astype(np.longdouble), axis=axis).take(-1, axis=axis)


def test_long_double_is_x87():
    assert np.finfo(np.longdouble).nmant == 63  # the independent check below needs 80-bit long double


def test_row_means_and_scale_match_numpy_longdouble(oracle):
    rng = np.random.default_rng(3)
    m = rng.gamma(2.0, 1.5, (257, 91)) * (rng.random((257, 91)) > 0.3)
    want = (_ld_seq_sum(m, 1) / np.longdouble(91)).astype(np.float64)
    assert np.array_equal(oracle.row_means(m), want)
    # scale(): colMeans, centre, sqrt(sum(v^2) / (n - 1)), divide — every double op rounded separately
    center = (_ld_seq_sum(m, 0) / np.longdouble(257)).astype(np.float64)
    v = m - center
    ss = _ld_seq_sum(v * v, 0).astype(np.float64)
    want = v / np.sqrt(ss / 256.0)
    assert np.array_equal(oracle.scale(m), want)
    # and they are what they claim to be, to rounding
    assert np.allclose(oracle.row_means(m), m.mean(1), rtol=1e-14)
    assert np.allclose(oracle.scale(m).std(0, ddof=1), 1.0, rtol=1e-12)


def test_mean_nonzero_is_r_mean_of_the_subset(oracle):
    rng = np.random.default_rng(4)
    m = rng.gamma(2.0, 1.5, (300, 40)) * (rng.random((300, 40)) > 0.5)
    nz = np.asfortranarray(m).ravel(order="F")
    nz = nz[nz != 0]
    assert oracle.mean_nonzero(m) == oracle.r_mean(nz)
    assert abs(oracle.mean_nonzero(m) - nz.mean()) < 1e-13


def test_set_gexp_mats_vignette_form_is_gexp_minus_ref(oracle, golden_inputs, golden_gexp):
    gexp, ref = golden_inputs["gexp"], golden_inputs["ref"]
    norm, keep = oracle.set_gexp_mats(gexp, ref, filter=False, scale=False)  # Getting_Started.Rmd:52
    assert np.array_equal(keep, np.arange(gexp.shape[0]))
    assert np.array_equal(norm, gexp - ref[:, None])
    # the committed expression fixture was made from the same matrices
    g32 = golden_gexp["gexp_norm_f32"]
    assert np.array_equal(norm[:, :40].astype(np.float32), g32[:800])


def test_set_gexp_mats_filter_and_scale_match_numpy(oracle, golden_inputs):
    gexp, ref = golden_inputs["gexp"], golden_inputs["ref"][:, None]
    norm, keep = oracle.set_gexp_mats(gexp, ref, filter=True, scale=True)
    rs, rr = oracle.row_means(gexp), oracle.row_means(ref)
    mt, mr = oracle.mean_nonzero(gexp), oracle.mean_nonzero(ref)
    want_keep = np.nonzero(((rs > 0) & (rr > 0)) | (rs > mt) | (rr > mr))[0]
    assert np.array_equal(keep, want_keep) and 0 < keep.size < gexp.shape[0]
    want = oracle.scale(gexp[keep]) - oracle.row_means(oracle.scale(ref[keep]))[:, None]
    assert np.array_equal(norm, want)


def test_set_allele_mats_reproduces_the_reference_printouts(oracle, golden_inputs, golden_allele):
    r, cov = golden_inputs["r"].astype(np.int32), golden_inputs["cov"].astype(np.int32)
    for thr, (printed, final) in zip((0.05, 0.1), golden_inputs["known_counts"]):
        out = oracle.set_allele_mats(r, cov, het_deviance_threshold=thr)
        l, nb = (r > 0).sum(1), (cov > 0).sum(1)
        with np.errstate(all="ignore"):
            E = l / nb
        assert int(((E > thr) & (E < 1 - thr)).sum()) == printed  # docs/Integrating.md:27, Getting_Started.md:230
        assert out["keep"].size == final
    # threshold 0.1 is what the committed allele fixture holds
    assert np.array_equal(out["r_maf"], golden_allele["r_maf"].astype(np.int32))
    assert np.array_equal(out["n_sc"], golden_allele["n_sc"].astype(np.int32))
    assert np.array_equal(out["l_maf"], golden_allele["l_maf"])
    assert np.array_equal(out["n_bulk"], golden_allele["n_bulk"])


def test_set_allele_mats_without_filter_codes_na_rows_as_zero(oracle):
    rng = np.random.default_rng(6)
    n = rng.integers(0, 4, (50, 12)).astype(np.int32)
    n[7] = 0  # no coverage at all: E = 0/0 = NA -> r.maf row of zeros, l.maf 0
    r = rng.binomial(n, 0.6).astype(np.int32)
    out = oracle.set_allele_mats(r, n, filter=False)
    assert out["keep"].size == 50 and (out["r_maf"][7] == 0).all() and out["l_maf"][7] == 0
    l, nb = (r > 0).sum(1), (n > 0).sum(1)
    flip = np.divide(l, nb, out=np.zeros(50), where=nb > 0) > 0.5
    assert np.array_equal(out["r_maf"], np.where(flip[:, None], n - r, r))
    assert np.array_equal(out["l_maf"], np.where(flip, nb - l, l))


def test_retest_closed_forms_equal_brute_force_enumeration(oracle):
    """oracle.retest_gexp / retest_comb (the closed forms the kernels are checked against) vs summing the
    joint of inst/bug/expressionModel.bug / combinedModel.bug over every (S_1..S_K, dd) for K = 5 cells;
    alpha_k, beta ~ U(0,1) integrate to prior 1/2 on each S_k and on dd."""
    import itertools
    rng = np.random.default_rng(9)
    J, K, mag = 7, 5, 0.9
    x = rng.standard_normal((J, K))
    x[:, :2] -= mag
    prec = rng.uniform(0.5, 1.5, K)
    off = rng.normal(0, 3, K)  # per-cell allele log-odds of the combined model

    def ll(mu):
        return (0.5 * np.log(prec / (2 * np.pi)) - 0.5 * prec * (x - mu) ** 2).sum(0)

    l0, lp, lm = ll(0.0), ll(mag), ll(-mag)
    for extra in (None, off):
        amp, dele, Z = np.zeros(K), np.zeros(K), 0.0
        for dd in (0, 1):
            for S in itertools.product((0, 1), repeat=K):
                S = np.array(S)
                l1 = lp if dd else (lm + (0 if extra is None else extra))
                w = np.exp(np.where(S == 1, l1, l0).sum() - l0.sum())
                Z += w
                amp += w * S * dd
                dele += w * S * (1 - dd)
        pa, pd_ = oracle.retest_gexp(x, mag, prec, del_loglik=extra)
        assert np.abs(pa - amp / Z).max() < 1e-12 and np.abs(pd_ - dele / Z).max() < 1e-12


def _rle_values(seq):
    out = []
    for v in seq:
        if not out or out[-1] != v:
            out.append(int(v))
    return out


def test_stateless_allele_path_reproduces_the_vignette_figure(oracle, golden_inputs):
    """docs/figure/Integrating/unnamed-chunk-6-1.png: plotAlleleProfile(region = calcAlleleCnvBoundaries(...)
    $region) on the bundled data draws one panel per run of region@seqnames (R/HoneyBADGER_allele.R:248):
    chr1 chr3 chr9 chr10 chr12 chr13 chr14 chr15 chr19.  The restated path — setAlleleMats at the default
    threshold, runmean(k=31) -> Ward.D2 -> cutree(1..3), pooled counts, 2-state binomial Viterbi, vote,
    runs, trim, range() per chromosome — must land on the same chromosomes in the same order."""
    from oracle import pyref as P
    from oracle.oracle import runmean, ward_groups
    r, cov = golden_inputs["r"].astype(np.int32), golden_inputs["cov"].astype(np.int32)
    out = oracle.set_allele_mats(r, cov, het_deviance_threshold=0.05)
    assert out["keep"].size == 4550
    chrom = golden_inputs["chrom"][out["keep"]]
    r_maf, n_sc = out["r_maf"].astype(float), out["n_sc"].astype(float)
    with np.errstate(all="ignore"):
        labels = ward_groups(runmean(r_maf / n_sc, 31), [1, 2, 3])
    res = P.calc_allele_cnv_boundaries_fn(r_maf, n_sc, labels)
    region_chroms = [c for idx in res["info"] for c in _rle_values(chrom[idx])]
    assert _rle_values(region_chroms) == list(golden_inputs["figure_region_chroms"])
