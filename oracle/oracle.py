"""ctypes binding of oracle/libhb_oracle.so (the C restatement in hb_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhb_oracle.so")

HBO_OK = 0
HBO_ERR_UNDERFLOW = -3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        d, i32, i64, p = C.c_double, C.c_int, C.c_int64, C.c_void_p
        L.hbo_dnorm_log.restype = d
        L.hbo_dnorm_log.argtypes = [d, d, d]
        L.hbo_dnorm.restype = d
        L.hbo_dnorm.argtypes = [d, d, d]
        L.hbo_stirlerr.restype = d
        L.hbo_stirlerr.argtypes = [d]
        L.hbo_bd0.restype = d
        L.hbo_bd0.argtypes = [d, d]
        L.hbo_dbinom.restype = d
        L.hbo_dbinom.argtypes = [d, d, d, i32]
        L.hbo_r_mean.restype = d
        L.hbo_r_mean.argtypes = [p, i64]
        L.hbo_r_sd.restype = d
        L.hbo_r_sd.argtypes = [p, i64]
        L.hbo_viterbi_logb.restype = i32
        L.hbo_viterbi_logb.argtypes = [i32, i64, p, p, p, p, p]
        L.hbo_viterbi_norm.restype = i32
        L.hbo_viterbi_norm.argtypes = [i32, i64, p, p, p, p, p, p]
        L.hbo_viterbi_binom.restype = i32
        L.hbo_viterbi_binom.argtypes = [i32, i64, p, p, p, p, p, p]
        L.hbo_forwardback_prob.restype = i32
        L.hbo_forwardback_prob.argtypes = [i32, i64, p, p, p, p, p, p, p]
        L.hbo_forwardback_norm.restype = i32
        L.hbo_forwardback_norm.argtypes = [i32, i64, p, p, p, p, p, p, p]
        L.hbo_forwardback_binom.restype = i32
        L.hbo_forwardback_binom.argtypes = [i32, i64, p, p, p, p, p, p, p]
        L.hbo_fbv_norm_batch.restype = i32
        L.hbo_fbv_norm_batch.argtypes = [i64, i64, i64, p, p, i32, p, p, i32, p, p, p, p, p, i32]
        L.hbo_fbv_binom_batch.restype = i32
        L.hbo_fbv_binom_batch.argtypes = [i64, i64, i64, p, p, p, i32, p, p, p, p, p, p, i32]
        L.hbo_fbv_binom_batch_rho.restype = i32
        L.hbo_fbv_binom_batch_rho.argtypes = [i64, i64, i64, p, p, p, i32, p, d, p, p, p, p, p, i32]
        L.hbo_dbetabinom_log.restype = d
        L.hbo_dbetabinom_log.argtypes = [d, d, d, d]
        L.hbo_pool_mean.restype = None
        L.hbo_pool_mean.argtypes = [p, i64, p, i64, p]
        L.hbo_pool_count_pos.restype = None
        L.hbo_pool_count_pos.argtypes = [p, i64, p, i64, p]
        L.hbo_row_means.restype = None
        L.hbo_row_means.argtypes = [p, i64, i64, p]
        L.hbo_scale.restype = None
        L.hbo_scale.argtypes = [p, i64, i64]
        L.hbo_mean_nonzero.restype = d
        L.hbo_mean_nonzero.argtypes = [p, i64]
        L.hbo_set_gexp_mats.restype = i64
        L.hbo_set_gexp_mats.argtypes = [p, i64, i64, p, i64, i32, d, p, p, i32, p, p]
        L.hbo_set_allele_mats.restype = i64
        L.hbo_set_allele_mats.argtypes = [p, p, i64, i64, p, p, i32, d, i32, p, p, p, p, p, p]
        L.hbo_num_threads.restype = i32
        _lib = L
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------- scalar densities
def dnorm_log(x, mu, sigma):
    return lib().hbo_dnorm_log(float(x), float(mu), float(sigma))


def dnorm(x, mu, sigma):
    return lib().hbo_dnorm(float(x), float(mu), float(sigma))


def dbinom(x, n, p, log=False):
    return lib().hbo_dbinom(float(x), float(n), float(p), int(bool(log)))


def stirlerr(n):
    return lib().hbo_stirlerr(float(n))


def bd0(x, np_):
    return lib().hbo_bd0(float(x), float(np_))


def r_mean(x):
    x = _f64(x)
    return lib().hbo_r_mean(_p(x), x.size)


def r_sd(x):
    x = _f64(x)
    return lib().hbo_r_sd(_p(x), x.size)


# ---------------------------------------------------------------- single chains
def sym_Pi(S: int, t: float) -> np.ndarray:
    """Transition matrix of R/HoneyBADGER.R:1150 (S=3: 1-2t / t) and :1409 (S=2: 1-t / t)."""
    Pi = np.full((S, S), t, dtype=np.float64)
    np.fill_diagonal(Pi, 1 - (S - 1) * t)
    return Pi


def viterbi_norm(x, mean, sd, Pi, delta):
    """HiddenMarkov::Viterbi(dthmm(x, Pi, delta, "norm", list(mean, sd))) -> 1-based states."""
    x = _f64(x)
    mean, sd, Pi, delta = _f64(mean), _f64(sd), _f64(Pi), _f64(delta)
    S = mean.size
    y = np.zeros(x.size, dtype=np.int32)
    rc = lib().hbo_viterbi_norm(S, x.size, _p(x), _p(mean), _p(sd), _p(Pi), _p(delta), _p(y))
    if rc == HBO_ERR_UNDERFLOW:
        raise FloatingPointError("Problems With Underflow")
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return y


def viterbi_binom(x, size, prob, Pi, delta):
    x, size = _i32(x), _i32(size)
    prob, Pi, delta = _f64(prob), _f64(Pi), _f64(delta)
    S = prob.size
    y = np.zeros(x.size, dtype=np.int32)
    rc = lib().hbo_viterbi_binom(S, x.size, _p(x), _p(size), _p(prob), _p(Pi), _p(delta), _p(y))
    if rc == HBO_ERR_UNDERFLOW:
        raise FloatingPointError("Problems With Underflow")
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return y


def viterbi_logb(logb, Pi, delta, return_nu=False):
    logb = _f64(logb)
    T, S = logb.shape
    Pi, delta = _f64(Pi), _f64(delta)
    y = np.zeros(T, dtype=np.int32)
    nu = np.zeros((T, S)) if return_nu else None
    rc = lib().hbo_viterbi_logb(S, T, _p(logb), _p(Pi), _p(delta), _p(y), _p(nu))
    if rc == HBO_ERR_UNDERFLOW:
        raise FloatingPointError("Problems With Underflow")
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return (y, nu) if return_nu else y


def forwardback_norm(x, mean, sd, Pi, delta):
    """(gamma[T,S], LL) with gamma = exp(logalpha + logbeta - LL) of HiddenMarkov::forwardback."""
    x = _f64(x)
    mean, sd, Pi, delta = _f64(mean), _f64(sd), _f64(Pi), _f64(delta)
    S = mean.size
    g = np.zeros((x.size, S))
    ll = C.c_double(0)
    rc = lib().hbo_forwardback_norm(S, x.size, _p(x), _p(mean), _p(sd), _p(Pi), _p(delta), _p(g),
                                    C.addressof(ll))
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return g, ll.value


def forwardback_binom(x, size, prob, Pi, delta):
    x, size = _i32(x), _i32(size)
    prob, Pi, delta = _f64(prob), _f64(Pi), _f64(delta)
    S = prob.size
    g = np.zeros((x.size, S))
    ll = C.c_double(0)
    rc = lib().hbo_forwardback_binom(S, x.size, _p(x), _p(size), _p(prob), _p(Pi), _p(delta),
                                     _p(g), C.addressof(ll))
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return g, ll.value


def forwardback_prob(prob, Pi, delta):
    prob = _f64(prob)
    T, S = prob.shape
    Pi, delta = _f64(Pi), _f64(delta)
    la, lb, g = np.zeros((T, S)), np.zeros((T, S)), np.zeros((T, S))
    ll = C.c_double(0)
    rc = lib().hbo_forwardback_prob(S, T, _p(prob), _p(Pi), _p(delta), _p(la), _p(lb), _p(g),
                                    C.addressof(ll))
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    return la, lb, g, ll.value


# ---------------------------------------------------------------- gene-major batches
def fbv_norm_batch(x, seg_off, mean, sd, Pi, delta, want_gamma=True, nthreads=0):
    """x[T,B] gene-major; one chain per (column, segment).  sd is [B] or [nseg,B].

    Returns (y uint8 [T,B] 1-based (0 = chain hit underflow), gamma [3,T,B] | None, ll [nseg,B], rc).
    """
    x = _f64(x)
    T, B = x.shape
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    nseg = seg_off.size - 1
    sd = _f64(sd)
    per_seg = int(sd.ndim == 2)
    mean, Pi, delta = _f64(mean), _f64(Pi), _f64(delta)
    y = np.zeros((T, B), dtype=np.uint8)
    g = np.zeros((3, T, B)) if want_gamma else None
    ll = np.zeros((nseg, B))
    nt = nthreads or lib().hbo_num_threads()
    rc = lib().hbo_fbv_norm_batch(T, B, B, _p(x), _p(seg_off), nseg, _p(mean), _p(sd), per_seg,
                                  _p(Pi), _p(delta), _p(y), _p(g), _p(ll), nt)
    return y, g, ll, rc


def dbetabinom_log(x, n, p, rho):
    """log beta-binomial pmf (mean p, intra-class correlation rho): binomial + correction, see hb_oracle.c"""
    return lib().hbo_dbetabinom_log(float(x), float(n), float(p), float(rho))


def fbv_binom_batch(x, n, seg_off, prob, Pi, delta, want_gamma=True, nthreads=0, rho=0.0):
    x, n = _i32(x), _i32(n)
    T, B = x.shape
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    nseg = seg_off.size - 1
    prob, Pi, delta = _f64(prob), _f64(Pi), _f64(delta)
    y = np.zeros((T, B), dtype=np.uint8)
    g = np.zeros((2, T, B)) if want_gamma else None
    ll = np.zeros((nseg, B))
    nt = nthreads or lib().hbo_num_threads()
    rc = lib().hbo_fbv_binom_batch_rho(T, B, B, _p(x), _p(n), _p(seg_off), nseg, _p(prob), float(rho), _p(Pi),
                                       _p(delta), _p(y), _p(g), _p(ll), nt)
    return y, g, ll, rc


# ---------------------------------------------------------------- pooled statistics
def pool_mean(mat_gn, cols):
    """apply(M[, cols], 1, mean) for M given as a numpy [G, N] array (any memory order)."""
    m = np.asfortranarray(mat_gn, dtype=np.float64)
    G = m.shape[0]
    cols = _i32(cols)
    out = np.zeros(G)
    lib().hbo_pool_mean(m.ctypes.data_as(C.c_void_p), G, _p(cols), cols.size, _p(out))
    return out


def pool_count_pos(mat_gn, cols):
    """rowSums(M[, cols] > 0)."""
    m = np.asfortranarray(mat_gn, dtype=np.float64)
    G = m.shape[0]
    cols = _i32(cols)
    out = np.zeros(G, dtype=np.int32)
    lib().hbo_pool_count_pos(m.ctypes.data_as(C.c_void_p), G, _p(cols), cols.size, _p(out))
    return out


# ---------------------------------------------------------------- input construction (8f-3)
def row_means(mat_gn):
    """rowMeans(M)."""
    m = np.asfortranarray(mat_gn, dtype=np.float64)
    out = np.zeros(m.shape[0])
    lib().hbo_row_means(_p(m), m.shape[0], m.shape[1], _p(out))
    return out


def scale(mat_gn):
    """scale(M) (centre and scale every column)."""
    m = np.array(mat_gn, dtype=np.float64, order="F", copy=True)
    lib().hbo_scale(_p(m), m.shape[0], m.shape[1])
    return m


def mean_nonzero(mat_gn):
    """mean(M[M != 0])."""
    m = np.asfortranarray(mat_gn, dtype=np.float64)
    return float(lib().hbo_mean_nonzero(_p(m), m.size))


def set_gexp_mats(sc, ref, filter=True, min_mean_both=0.0, min_mean_test=None, min_mean_ref=None, scale=True):
    """Arithmetic of setGexpMats (R/HoneyBADGER.R:124-161) for row-aligned sc [G, N], ref [G, R].
    Returns (gexp.norm [Gk, N], kept row indices)."""
    sc = np.asfortranarray(sc, dtype=np.float64)
    ref = np.asfortranarray(np.asarray(ref, dtype=np.float64).reshape(sc.shape[0], -1))
    G, N = sc.shape
    norm = np.zeros(G * N)
    keep = np.zeros(G, dtype=np.int32)
    mt = None if min_mean_test is None else C.c_double(min_mean_test)
    mr = None if min_mean_ref is None else C.c_double(min_mean_ref)
    Gk = lib().hbo_set_gexp_mats(_p(sc), G, N, _p(ref), ref.shape[1], int(filter), float(min_mean_both),
                                 None if mt is None else C.byref(mt), None if mr is None else C.byref(mr),
                                 int(scale), _p(norm), _p(keep))
    return norm[:Gk * N].reshape(N, Gk).T, keep[:Gk]


def set_allele_mats(r, n, l_init=None, n_bulk_init=None, filter=True, het_deviance_threshold=0.05, min_cell=3):
    """Arithmetic of setAlleleMats (R/HoneyBADGER.R:486-595) for integer matrices r, n [G, N].
    Returns dict(r_maf, n_sc, l, n_bulk, l_maf, keep)."""
    r = np.asfortranarray(r, dtype=np.int32)
    n = np.asfortranarray(n, dtype=np.int32)
    G, N = r.shape
    rm, ns = np.zeros(G * N, np.int32), np.zeros(G * N, np.int32)
    l, nb, lm = np.zeros(G), np.zeros(G), np.zeros(G)
    keep = np.zeros(G, dtype=np.int32)
    li = None if l_init is None else _f64(l_init)
    ni = None if n_bulk_init is None else _f64(n_bulk_init)
    Gk = lib().hbo_set_allele_mats(_p(r), _p(n), G, N, _p(li), _p(ni), int(filter), float(het_deviance_threshold),
                                   int(min_cell), _p(rm), _p(ns), _p(l), _p(nb), _p(lm), _p(keep))
    return dict(r_maf=rm[:Gk * N].reshape(N, Gk).T, n_sc=ns[:Gk * N].reshape(N, Gk).T, l=l[:Gk],
                n_bulk=nb[:Gk], l_maf=lm[:Gk], keep=keep[:Gk])


def retest_gexp(x_rows, mag0, sigma0=None, del_loglik=None):
    """Closed-form posterior of inst/bug/expressionModel.bug, plain numpy, written
    from the likelihoods (not from the kernel's difference form): returns (P(amp), P(del)) per cell.
    x_rows: [genes of the region, cells]; sigma0: per-cell dnorm PRECISION as the reference passes it.
    del_loglik: per-cell log-likelihood of further data under "deleted" minus under "not deleted"
    (inst/bug/combinedModel.bug: the allele counts), multiplying the deletion branch only."""
    x = np.asarray(x_rows, dtype=np.float64)
    n, N = x.shape
    prec = np.ones(N) if sigma0 is None else np.asarray(sigma0, dtype=np.float64)

    def ll(mu):
        return (0.5 * np.log(prec / (2 * np.pi)) - 0.5 * prec * (x - mu) ** 2).sum(0)

    l0, lp, lm = ll(0.0), ll(mag0), ll(-mag0)
    if del_loglik is not None:
        lm = lm + np.asarray(del_loglik, dtype=np.float64)
    La = np.logaddexp(l0, lp) + np.log(0.5)
    Ld = np.logaddexp(l0, lm) + np.log(0.5)
    D = La.sum() - Ld.sum()
    post1 = 1.0 / (1.0 + np.exp(-D)) if D >= 0 else np.exp(D) / (1.0 + np.exp(D))
    return post1 * np.exp(lp + np.log(0.5) - La), (1 - post1) * np.exp(lm + np.log(0.5) - Ld)


def retest_allele_model(r_rows, n_rows, l_maf, n_bulk, gene=None, pseudo=0.1, mono=0.7, logodds=False):
    """Exact per-cell marginal P(S_k = 1 | data) of inst/bug/snpModel.bug (sampled with JAGS by
    calcAlleleCnvProb, R/HoneyBADGER.R:907-1042) with the shared allele indicators ma at their bulk
    posterior — the quantity libhb_b200's hb_retest_allele_model_dev computes.  Plain numpy; the integral
    over h ~ dnorm(0.5, 0.1)T(pseudo, 1 - pseudo) by adaptive quadrature (scipy.integrate.quad)."""
    from scipy import integrate, special
    r = np.asarray(r_rows, dtype=np.float64)
    n = np.asarray(n_rows, dtype=np.float64)
    m, N = r.shape
    gene = np.arange(m) if gene is None else np.asarray(gene)
    sd = 1.0 / np.sqrt(0.1)
    lo, hi = pseudo, 1.0 - pseudo
    Z = integrate.quad(lambda h: np.exp(-0.5 * ((h - 0.5) / sd) ** 2), lo, hi, epsabs=0, epsrel=1e-13)[0]
    lq_odds = (np.asarray(n_bulk, float) - 2.0 * np.asarray(l_maf, float)) * np.log((1 - pseudo) / pseudo)
    lq1, lq0 = -np.logaddexp(0.0, -lq_odds), -np.logaddexp(0.0, lq_odds)
    lp, l1p = np.log(pseudo), np.log1p(-pseudo)
    cache = {}

    def log_int(rv, nv):
        key = (rv, nv)
        if key not in cache:
            pk = rv / nv  # integrand peak: scale it out so quad sees O(1) values
            pk = min(max(pk, lo), hi)
            ref = rv * np.log(pk) + (nv - rv) * np.log1p(-pk)
            f = lambda h: np.exp(rv * np.log(h) + (nv - rv) * np.log1p(-h) - ref - 0.5 * ((h - 0.5) / sd) ** 2)
            val = integrate.quad(f, lo, hi, points=[pk] if lo < pk < hi else None, epsabs=0, epsrel=1e-12,
                                 limit=400)[0]
            cache[key] = ref + np.log(val / Z)
        return cache[key]

    out = np.zeros(N)
    uniq = list(dict.fromkeys(gene.tolist()))
    for c in range(N):
        L1 = L0 = 0.0
        for g in uniq:
            A = B1 = B0 = 0.0
            for j in np.nonzero(gene == g)[0]:
                if n[j, c] > 0:
                    a = r[j, c] * lp + (n[j, c] - r[j, c]) * l1p
                    b = r[j, c] * l1p + (n[j, c] - r[j, c]) * lp
                    L1 += np.logaddexp(lq1[j] + a, lq0[j] + b)
                    B1 += a
                    B0 += b
                    A += log_int(r[j, c], n[j, c])
            L0 += np.logaddexp(np.log1p(-mono) + A, np.log(mono / 2) + np.logaddexp(B1, B0))
        out[c] = (L1 - L0) if logodds else special.expit(L1 - L0)
    return out


def retest_comb(x_rows, mag0, r_rows, n_rows, l_maf, n_bulk, sigma0=None, gene=None, pseudo=0.1, mono=0.7):
    """Marginal posterior of inst/bug/combinedModel.bug: (P(amp), P(del)) per cell."""
    off = retest_allele_model(r_rows, n_rows, l_maf, n_bulk, gene=gene, pseudo=pseudo, mono=mono, logodds=True)
    return retest_gexp(x_rows, mag0, sigma0, del_loglik=off)


# --------------------------------------------------------------------------- grouping (checker for hb_group.cuh)
# numpy / scipy restatements of caTools::runmean and cutree(hclust(dist(.), "ward.D2")) — test infrastructure
# for the device grouping kernels; they lived in the package until round 2.
def runmean(mat: np.ndarray, k: int) -> np.ndarray:
    """caTools::runmean(x, k) along rows of each column: centred window, shrinking at the ends,
    non-finite values skipped (alg="C", endrule="mean")."""
    G = mat.shape[0]
    ok = np.isfinite(mat)
    v = np.where(ok, mat, 0.0)
    zero = np.zeros((1,) + mat.shape[1:])
    cs = np.vstack([zero, np.cumsum(v, 0)])
    cn = np.vstack([zero, np.cumsum(ok, 0)])
    h = k // 2
    lo = np.clip(np.arange(G) - h, 0, G)
    hi = np.clip(np.arange(G) + h + 1, 0, G)
    with np.errstate(all="ignore"):
        return (cs[hi] - cs[lo]) / (cn[hi] - cn[lo])


def ward_groups(mat_smooth: np.ndarray, heights) -> list:
    """cutree(hclust(dist(t(mat.smooth)), "ward.D2"), k=h) for each h; clusters numbered by first
    appearance in cell order (so group 1 always holds the first cell)."""
    from scipy.cluster.hierarchy import cut_tree, linkage
    X = np.nan_to_num(mat_smooth.T, nan=0.0, posinf=0.0, neginf=0.0)
    N = X.shape[0]
    if N == 1:
        return [np.ones(1, dtype=np.int32) for _ in heights]
    Z = linkage(X, method="ward")
    out = []
    for h in heights:
        lab = cut_tree(Z, n_clusters=int(h))[:, 0]
        remap, res = {}, np.zeros(N, dtype=np.int32)
        for i, v in enumerate(lab):
            res[i] = remap.setdefault(v, len(remap) + 1)
        out.append(res)
    return out
