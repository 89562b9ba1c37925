"""Host-side mirror of HoneyBADGER's interface for the HMM segmentation path.

Same names, argument meaning, defaults and quirks as the reference for

    HoneyBADGER$calcGexpCnvBoundaries   R/HoneyBADGER.R:1091-1316   (recursive, stateful)
    HoneyBADGER$calcAlleleCnvBoundaries R/HoneyBADGER.R:1336-1578
    calcGexpCnvBoundaries               R/HoneyBADGER_gexp.R:462-584  (one round, stateless)
    calcAlleleCnvBoundaries             R/HoneyBADGER_allele.R:543-651

The matrices are uploaded once (gene-major, fp64 / int32) and stay on the GPU; each recursion
level sends only the candidate cell sets (a [K x N] 0/1 membership) and gets back the K Viterbi
paths.  Pooling (R `mean`/`sd`/`rowSums(.>0)` semantics), emissions and the Viterbi recursion run
in libhb_b200.so; what stays in Python is the bookkeeping the reference does with names, votes and
runs (vectorised here; the literal loops live in oracle/pyref.py for the tests).

Outside the path (SURVEY.md §8f "next" rows), supplied as replaceable hooks:
  * grouping — `dist` + `hclust(ward.D2)` + `cutree` on the windowed matrix (R/HoneyBADGER.R:1125-1136):
    `ward_groups_dev` does it on the device (runmean, R dist with its NA rule, Ward.D2 by nearest-neighbour
    chain); any f(mat, heights) can be plugged in (the numpy / scipy restatement the tests
    check it against lives in oracle/oracle.py — the package itself has no CPU compute).
  * retest   — the JAGS posteriors (`calcGexpCnvProb`, `calcAlleleCnvProb`): `retest_gexp_closed_form`
    (exact marginal of inst/bug/expressionModel.bug) and `retest_allele_model` (per-cell
    marginal of inst/bug/snpModel.bug with the shared allele indicators at their bulk posterior), both
    CUDA kernels behind the C ABI; `retest_allele_lr` is a plain likelihood-ratio alternative.  None of
    them is parity-pinned: the reference samples these posteriors by MCMC.
R matrices with dimnames are pandas DataFrames here (index = gene / SNP names, columns = cells).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import pandas as pd

from . import _lib
from .hmm import sym_Pi

CHRS_DEFAULT = ["chr%d" % i for i in range(1, 23)]


# --------------------------------------------------------------------------- GRanges stand-in
@dataclass
class GRanges:
    """The slice of GenomicRanges::GRanges this path touches: names -> (seqnames, start, end)."""
    seqnames: np.ndarray
    start: np.ndarray
    end: np.ndarray
    names: list

    def __post_init__(self):
        self.seqnames = np.asarray(self.seqnames, dtype=object)
        self.start = np.asarray(self.start, dtype=np.float64)
        self.end = np.asarray(self.end, dtype=np.float64)
        self.names = list(self.names)
        self._ix = {n: i for i, n in enumerate(self.names)}

    def __getitem__(self, names):
        idx = [self._ix[n] for n in names]
        return GRanges(self.seqnames[idx], self.start[idx], self.end[idx], [self.names[i] for i in idx])

    def range(self):
        """range(gr): one (seqname, min start, max end) per chromosome present, in first-seen order."""
        out, seen = [], []
        for s in self.seqnames:
            if s not in seen:
                seen.append(s)
        for s in seen:
            m = self.seqnames == s
            out.append((s, float(self.start[m].min()), float(self.end[m].max())))
        return out


def snps_from_names(names) -> GRanges:
    """R/HoneyBADGER.R:598-608: 'chr:pos' / 'chr:start-end' row names -> GRanges ('chr' prefixed)."""
    chrom, st, en = [], [], []
    for n in names:
        parts = str(n).replace("-", ":").replace(" ", ":").split(":")
        c = parts[0] if parts[0].startswith("chr") else "chr" + parts[0]
        chrom.append(c)
        st.append(float(parts[1]))
        en.append(float(parts[2]) if len(parts) > 2 else float(parts[1]))
    return GRanges(chrom, st, en, list(names))


def _as_frame(m, prefix_r="g", prefix_c="cell") -> pd.DataFrame:
    if isinstance(m, pd.DataFrame):
        return m
    a = np.asarray(m, dtype=np.float64)
    return pd.DataFrame(a, index=[f"{prefix_r}{i + 1}" for i in range(a.shape[0])],
                        columns=[f"{prefix_c}{i + 1}" for i in range(a.shape[1])])


# --------------------------------------------------------------------------- hooks outside the path
def cut_merges(merge_a, merge_b, height, N, ks):
    """cutree(hc, k) for each k in `ks` from a merge list (leaf representatives, any order in which
    children precede parents): the N - k lowest merges are applied with union-find; clusters are
    numbered by first appearance in cell order, as cutree does (libhb_b200: hb_cutree_host)."""
    ma = np.ascontiguousarray(merge_a, dtype=np.int32)
    mb = np.ascontiguousarray(merge_b, dtype=np.int32)
    hh = np.ascontiguousarray(height, dtype=np.float64)
    kk = np.ascontiguousarray([int(k) for k in ks], dtype=np.int32)
    lab = np.empty((kk.size, int(N)), dtype=np.int32)
    _lib.check(_lib.load().hb_cutree_host(ma.ctypes.data, mb.ctypes.data, hh.ctypes.data, int(N), kk.ctypes.data,
                                          int(kk.size), lab.ctypes.data))
    return [lab[i].copy() for i in range(kk.size)]


def ward_groups_dev(x_dev, heights, k=101, counts=None):
    """Grouping on the device (libhb_b200: hb_runmean_dev / hb_dist_sq_dev / hb_ward_linkage_dev):
    cutree(hclust(dist(t(runmean(x, k))) with NA -> 0, "ward.D2"), h) for each h, as
    R/HoneyBADGER.R:1122-1136.  x_dev: gene-major fp64 CUDA tensor [G, N]; or counts=(r_dev, n_dev)
    int32 tensors for the allele form (r / n smoothed with k = 31, R/HoneyBADGER.R:1378-1385)."""
    import torch
    from .hmm import alloc_rows
    lib = _lib.load()
    if counts is not None:
        r, n = counts
        G, N = r.shape
        dev = r.device
    else:
        G, N = x_dev.shape
        dev = x_dev.device
    if N == 1:
        return [np.ones(1, dtype=np.int32) for _ in heights]
    S = alloc_rows(G, N, torch.float64, dev)
    if counts is not None:
        if r.stride(0) != n.stride(0):
            r, n = r.contiguous(), n.contiguous()
        _lib.check(lib.hb_runmean_ratio_dev(r.data_ptr(), n.data_ptr(), r.stride(0), G, N, int(k),
                                            S.data_ptr(), S.stride(0), _stream()))
    else:
        _lib.check(lib.hb_runmean_dev(x_dev.data_ptr(), x_dev.stride(0), G, N, int(k), S.data_ptr(),
                                      S.stride(0), _stream()))
    D = torch.empty((N, N), dtype=torch.float64, device=dev)
    _lib.check(lib.hb_dist_sq_dev(S.data_ptr(), S.stride(0), G, N, D.data_ptr(), _stream()))
    del S
    ma = torch.empty(N - 1, dtype=torch.int32, device=dev)
    mb = torch.empty(N - 1, dtype=torch.int32, device=dev)
    hh = torch.empty(N - 1, dtype=torch.float64, device=dev)
    _lib.check(lib.hb_ward_linkage_dev(D.data_ptr(), N, ma.data_ptr(), mb.data_ptr(), hh.data_ptr(), _stream()))
    return cut_merges(ma.cpu().numpy(), mb.cpu().numpy(), hh.cpu().numpy(), N, heights)


def retest_gexp_closed_form(gexp_rows, mag0, sigma0=None):
    """(P(amp), P(del)) per cell for one candidate region: the exact marginal posterior of
    inst/bug/expressionModel.bug, computed by libhb_b200's hb_retest_gexp_dev on the
    device.  JAGS' dnorm takes a PRECISION and the reference passes sigma0 there
    (expressionModel.bug:8) — kept.  Deterministic stand-in for the MCMC of calcGexpCnvProb
    (R/HoneyBADGER.R:374-470); agrees with it statistically, not bit for bit.
    gexp_rows: [genes of the region, cells] (CUDA fp64 tensor, or anything numpy can read)."""
    import torch
    x = gexp_rows if torch.is_tensor(gexp_rows) else \
        torch.as_tensor(np.ascontiguousarray(gexp_rows, dtype=np.float64), device="cuda")
    if x.stride(1) != 1:
        x = x.contiguous()
    n, N = x.shape
    rows = torch.arange(n, dtype=torch.int64, device=x.device)
    prec = None if sigma0 is None else torch.as_tensor(sigma0, dtype=torch.float64, device=x.device).contiguous()
    pa = torch.empty(N, dtype=torch.float64, device=x.device)
    pd_ = torch.empty(N, dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().hb_retest_gexp_dev(
        x.data_ptr(), x.stride(0), n, N, rows.data_ptr(), n, float(mag0),
        None if prec is None else prec.data_ptr(), pa.data_ptr(), pd_.data_ptr(), None, None, None, 0,
        _stream()))
    return pa.cpu().numpy(), pd_.cpu().numpy()


def retest_allele_lr(r_rows, n_rows, pd_=0.1, pn=0.45):
    """P(deletion) per cell for one candidate region from the binomial likelihood ratio pooled over
    the region's SNPs (equal priors), libhb_b200's hb_retest_allele_lr_dev.  Stand-in for
    calcAlleleCnvProb's JAGS snpModel (R/HoneyBADGER.R:907-1042), not parity-pinned."""
    import torch

    def dev_i32(a):
        t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(np.rint(a), dtype=np.int32),
                                                          device="cuda")
        return t if t.stride(1) == 1 else t.contiguous()

    r, n = dev_i32(r_rows), dev_i32(n_rows)
    if r.stride(0) != n.stride(0):
        r, n = r.contiguous(), n.contiguous()
    m, N = r.shape
    rows = torch.arange(m, dtype=torch.int64, device=r.device)
    out = torch.empty(N, dtype=torch.float64, device=r.device)
    _lib.check(_lib.load().hb_retest_allele_lr_dev(r.data_ptr(), n.data_ptr(), r.stride(0), m, N,
                                                   rows.data_ptr(), m, float(pd_), float(pn),
                                                   out.data_ptr(), _stream()))
    return out.cpu().numpy()


def retest_allele_model(r_rows, n_rows, l_maf, n_bulk, gene=None, pe=0.1, mono=0.7, logodds=False,
                        keep_on_device=False):
    """P(deletion) per cell for one candidate region: the exact per-cell marginal of
    inst/bug/snpModel.bug (calcAlleleCnvProb, R/HoneyBADGER.R:907-1042, samples it with JAGS; the caller
    at :1486 passes pe = pd), libhb_b200's hb_retest_allele_model_dev.  `gene`: the geneFactor codes of
    the region's SNPs (None = one gene per SNP).  Statistical stand-in for the MCMC, not parity-pinned."""
    import torch

    def dev_i32(a):
        t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(np.rint(a), dtype=np.int32),
                                                          device="cuda")
        return t if t.stride(1) == 1 else t.contiguous()

    r, n = dev_i32(r_rows), dev_i32(n_rows)
    if r.stride(0) != n.stride(0):
        r, n = r.contiguous(), n.contiguous()
    m, N = r.shape
    order = np.arange(m)
    gene_end = None
    if gene is not None:  # group by first appearance, like unique(geneFactor) (R/HoneyBADGER.R:963)
        codes = pd.factorize(np.asarray(gene, dtype=object))[0]
        order = np.argsort(codes, kind="stable")
        sc = codes[order]
        gene_end = np.ascontiguousarray(np.r_[sc[1:] != sc[:-1], True], dtype=np.uint8)
    rows = torch.as_tensor(order, dtype=torch.int64, device=r.device)
    l2 = np.ascontiguousarray(np.asarray(l_maf, dtype=np.float64)[order])
    nb2 = np.ascontiguousarray(np.asarray(n_bulk, dtype=np.float64)[order])
    out = torch.empty(N, dtype=torch.float64, device=r.device)
    _lib.check(_lib.load().hb_retest_allele_model_dev(r.data_ptr(), n.data_ptr(), r.stride(0), m, N,
                                                      rows.data_ptr(), m, _vp(gene_end), _vp(l2), _vp(nb2),
                                                      float(pe), float(mono), int(bool(logodds)), out.data_ptr(),
                                                      _stream()))
    return out if keep_on_device else out.cpu().numpy()


def retest_comb_closed_form(gexp_rows, mag0, r_rows, n_rows, l_maf, n_bulk, sigma0=None, gene=None, pe=0.1,
                            mono=0.7):
    """(P(amp), P(del)) per cell for one region from expression AND allele data: the marginal posterior of
    inst/bug/combinedModel.bug (calcCombCnvProb, R/HoneyBADGER.R:1620-1782, samples it with JAGS).  The
    two data types share S_k and the direction dd, so it is the expression retest with every cell's
    deletion hypothesis shifted by that cell's allele log-odds (hb_retest_comb_dev)."""
    import torch
    off = retest_allele_model(r_rows, n_rows, l_maf, n_bulk, gene=gene, pe=pe, mono=mono, logodds=True,
                              keep_on_device=True)
    x = gexp_rows if torch.is_tensor(gexp_rows) else \
        torch.as_tensor(np.ascontiguousarray(gexp_rows, dtype=np.float64), device="cuda")
    if x.stride(1) != 1:
        x = x.contiguous()
    n, N = x.shape
    rows = torch.arange(n, dtype=torch.int64, device=x.device)
    prec = None if sigma0 is None else torch.as_tensor(sigma0, dtype=torch.float64, device=x.device).contiguous()
    pa = torch.empty(N, dtype=torch.float64, device=x.device)
    pd_ = torch.empty(N, dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().hb_retest_comb_dev(
        x.data_ptr(), x.stride(0), n, N, rows.data_ptr(), n, float(mag0),
        None if prec is None else prec.data_ptr(), off.data_ptr(), pa.data_ptr(), pd_.data_ptr(), None, None,
        None, 0, _stream()))
    return pa.cpu().numpy(), pd_.cpu().numpy()


# --------------------------------------------------------------------------- vote -> runs -> trim
def _round_half_even(v: float) -> int:
    return int(round(v))


def runs_from_vote(vote, min_num: int, trim: float | None):
    """R/HoneyBADGER.R:1178-1210 (+ trim :1476, _gexp.R:568, _allele.R:642), vectorised.

    Returns a list of index arrays (one per surviving run, genomic order) or None.  The first
    element of every run of voted positions is NOT part of the run (the reference's loop starts
    labelling at the second element); runs are kept when the labelled count >= min_num; `trim`
    keeps names[1:round(L - L*trim)] (half-even rounding, trailing end only).
    """
    vote = np.asarray(vote)
    if vote.size == 0 or vote.max() == 0:
        return None
    v = vote > 0
    lab = np.zeros(v.size, dtype=bool)
    lab[1:] = v[1:] & v[:-1]
    if not lab.any():
        return None
    d = np.diff(np.concatenate([[0], lab.view(np.int8), [0]]))
    starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
    out = []
    for a, b in zip(starts, ends):
        L = b - a
        if L < min_num:
            continue
        idx = np.arange(a, b)
        if trim is not None:
            r = _round_half_even(L - L * trim)
            if r < 0:
                raise IndexError("only 0's may be mixed with negative subscripts")
            idx = idx[:max(r, 1)]  # R: x[1:0] is x[c(1, 0)] = x[1]
        out.append(idx)
    return out or None


def group_members(labels):
    """[(group id, member index array)] in `unique(ct)` order (first appearance)."""
    labels = np.asarray(labels)
    _, first = np.unique(labels, return_index=True)
    order = labels[np.sort(first)]
    return [(g, np.nonzero(labels == g)[0]) for g in order]


def _membership(labels_per_height, N):
    """Stack the cell sets with > 1 member over (height, group) -> uint8 [K, N] + (h, g index)."""
    rows, tags = [], []
    for h, labels in enumerate(labels_per_height):
        for gi, (_, cells) in enumerate(group_members(labels)):
            if cells.size > 1:
                m = np.zeros(N, dtype=np.uint8)
                m[cells] = 1
                rows.append(m)
                tags.append((h, gi))
    return (np.stack(rows) if rows else np.zeros((0, N), np.uint8)), tags


# --------------------------------------------------------------------------- device rounds
def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gexp_round_dev(x_dev, member, mean, t):
    """K pooled 3-state chains on a resident gene-major fp64 tensor: (y int32 [K,G], x [K,G], sd [K])."""
    lib = _lib.load()
    G, N = x_dev.shape
    member = np.ascontiguousarray(member, dtype=np.uint8)
    K = member.shape[0]
    y = np.zeros((K, G), dtype=np.int32)
    px = np.zeros((K, G), dtype=np.float64)
    psd = np.zeros(K, dtype=np.float64)
    Pi, delta = sym_Pi(3, t), np.array([0.0, 1.0, 0.0])
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    _lib.check(lib.hb_gexp_pooled_viterbi_dev(x_dev.data_ptr(), x_dev.stride(0), G, N, _vp(member), K,
                                              _vp(mean), _vp(Pi), _vp(delta), _vp(y), _vp(px),
                                              _vp(psd), _stream()))
    return y, px, psd


def allele_round_dev(r_dev, n_dev, member, pd_, pn, t):
    """K pooled 2-state chains on resident gene-major int32 tensors: (y [K,G], mafl, sizel)."""
    lib = _lib.load()
    G, N = r_dev.shape
    member = np.ascontiguousarray(member, dtype=np.uint8)
    K = member.shape[0]
    y = np.empty((K, G), dtype=np.int32)      # filled by the call (an error raises before they are used)
    mafl = np.empty((K, G), dtype=np.int32)
    sizel = np.empty((K, G), dtype=np.int32)
    Pi, delta, prob = sym_Pi(2, t), np.array([0.0, 1.0]), np.array([pd_, pn], dtype=np.float64)
    _lib.check(lib.hb_allele_pooled_viterbi_dev(r_dev.data_ptr(), n_dev.data_ptr(), r_dev.stride(0), G,
                                                N, _vp(member), K, _vp(prob), _vp(Pi), _vp(delta),
                                                _vp(y), _vp(mafl), _vp(sizel), _stream()))
    return y, mafl, sizel


def _to_dev_f64(a):
    """R numeric matrix -> CUDA float64 [G, N].  A large column-major source (a DataFrame's .values, an R
    matrix) goes up as contiguous column slabs and is transposed on the device instead of being gathered
    row by row on the host (0.9 s per 1.6 GB)."""
    import torch
    a = np.asarray(a)
    if (a.ndim == 2 and a.size >= (1 << 24) and a.dtype != object and a.flags.f_contiguous
            and not a.flags.c_contiguous):
        out = torch.empty(a.shape, dtype=torch.float64, device="cuda")
        at = a.T
        step = max(1, (1 << 28) // max(1, at.shape[1] * a.dtype.itemsize))
        for j in range(0, at.shape[0], step):
            out[:, j:j + step] = torch.from_numpy(at[j:j + step]).cuda().to(torch.float64).T
        return out
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda")


def _to_dev_i32(a):
    """Counts (R doubles, or any integer dtype) -> CUDA int32, rounded half-even like as.integer(rint(.)).
    Large matrices go up in row slabs in their own dtype and are converted on the device: no full-size
    host temporaries (a 200k x 10k matrix otherwise costs two 16 GB numpy copies before the upload)."""
    import torch
    a = np.asarray(a)
    if a.ndim != 2 or a.size < (1 << 24) or a.dtype == object:
        return torch.as_tensor(np.ascontiguousarray(np.rint(a), dtype=np.int32), device="cuda")
    out = torch.empty(a.shape, dtype=torch.int32, device="cuda")
    conv = lambda c: (c.round() if c.is_floating_point() else c).to(torch.int32)  # noqa: E731
    if a.flags.f_contiguous and not a.flags.c_contiguous:
        # a DataFrame's .values (pandas keeps the block column-major, like R): upload contiguous column
        # slabs and transpose on the device instead of gathering strided rows on the host
        at = a.T
        step = max(1, (1 << 28) // max(1, at.shape[1] * a.dtype.itemsize))
        for j in range(0, at.shape[0], step):
            out[:, j:j + step] = conv(torch.from_numpy(at[j:j + step]).cuda()).T
        return out
    step = max(1, (1 << 28) // max(1, a.shape[1] * a.dtype.itemsize))
    for i in range(0, a.shape[0], step):
        out[i:i + step] = conv(torch.from_numpy(np.ascontiguousarray(a[i:i + step])).cuda())
    return out


# --------------------------------------------------------------------------- input construction (8f-3)
def set_gexp_mats_dev(sc, ref, filter=True, minMeanBoth=0.0, minMeanTest=None, minMeanRef=None, scale=True):
    """Arithmetic of setGexpMats (R/HoneyBADGER.R:139-161) on the device: rowMeans filter, scale(),
    gexp.sc - rowMeans(gexp.ref), with R's 80-bit accumulators replayed.  sc [G, N], ref [G, R] are
    row-aligned numpy arrays or resident cuda fp64 tensors.  Returns (gexp.norm as a resident gene-major
    tensor [Gk, N], kept row indices, info bits)."""
    import torch
    lib = _lib.load()
    sc_d = sc if isinstance(sc, torch.Tensor) else _to_dev_f64(sc)
    ref_d = ref if isinstance(ref, torch.Tensor) else _to_dev_f64(np.asarray(ref, dtype=np.float64).reshape(sc_d.shape[0], -1))
    G, N = sc_d.shape
    R = ref_d.shape[1]
    norm = torch.empty((G, N), dtype=torch.float64, device="cuda")
    keep = np.zeros(G, dtype=np.int32)
    gk, info = C.c_int64(0), C.c_int32(0)
    mt = None if minMeanTest is None else C.byref(C.c_double(float(minMeanTest)))
    mr = None if minMeanRef is None else C.byref(C.c_double(float(minMeanRef)))
    _lib.check(lib.hb_set_gexp_mats_dev(sc_d.data_ptr(), sc_d.stride(0), G, N, ref_d.data_ptr(), ref_d.stride(0),
                                        R, int(bool(filter)), float(minMeanBoth), mt, mr, int(bool(scale)),
                                        norm.data_ptr(), norm.stride(0), _vp(keep), C.byref(gk), C.byref(info),
                                        _stream()))
    return norm[:gk.value], keep[:gk.value].copy(), info.value


def set_allele_mats_dev(r, n_sc, l_init=None, n_bulk_init=None, filter=True, het_deviance_threshold=0.05,
                        min_cell=3):
    """Arithmetic of setAlleleMats (R/HoneyBADGER.R:493-595) on the device: in-silico bulk, heterozygosity
    and coverage filters, lesser-allele flip.  r, n_sc [G, N]: numpy or resident cuda int32 tensors.
    Returns dict(r_maf, n_sc (resident [Gk, N]), l, n_bulk, l_maf (numpy [Gk]), keep, n_het)."""
    import torch
    lib = _lib.load()
    r_d = r if isinstance(r, torch.Tensor) else _to_dev_i32(r)
    n_d = n_sc if isinstance(n_sc, torch.Tensor) else _to_dev_i32(n_sc)
    G, N = r_d.shape
    o_r = torch.empty((G, N), dtype=torch.int32, device="cuda")
    o_n = torch.empty((G, N), dtype=torch.int32, device="cuda")
    l, nb, lm = np.zeros(G), np.zeros(G), np.zeros(G)
    keep = np.zeros(G, dtype=np.int32)
    gk, nhet = C.c_int64(0), C.c_int64(0)
    li = None if l_init is None or n_bulk_init is None else np.ascontiguousarray(l_init, dtype=np.float64)
    ni = None if li is None else np.ascontiguousarray(n_bulk_init, dtype=np.float64)
    _lib.check(lib.hb_set_allele_mats_dev(r_d.data_ptr(), n_d.data_ptr(), r_d.stride(0), G, N, _vp(li), _vp(ni),
                                          int(bool(filter)), float(het_deviance_threshold), int(min_cell),
                                          o_r.data_ptr(), o_n.data_ptr(), o_r.stride(0), _vp(l), _vp(nb),
                                          _vp(lm), _vp(keep), C.byref(gk), C.byref(nhet), _stream()))
    k = gk.value
    return dict(r_maf=o_r[:k], n_sc=o_n[:k], l=l[:k].copy(), n_bulk=nb[:k].copy(), l_maf=lm[:k].copy(),
                keep=keep[:k].copy(), n_het=nhet.value)


# --------------------------------------------------------------------------- ordering (a1)
def order_genes(rownames, genes: GRanges, chrs, has_na=None):
    """R/HoneyBADGER.R:1114-1121: per chromosome (in `chrs` order) sort by (start+end)/2 (stable),
    drop rows with NA, concatenate.  Returns positions into `rownames`."""
    g = genes[rownames]
    mid = (g.start + g.end) / 2
    keep = np.ones(len(rownames), dtype=bool) if has_na is None else ~np.asarray(has_na)
    out = []
    for c in chrs:
        ii = np.nonzero(g.seqnames == c)[0]
        if ii.size == 0:
            continue
        ii = ii[np.argsort(mid[ii], kind="stable")]
        out.append(ii[keep[ii]])
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


# --------------------------------------------------------------------------- stateless functions
def calcGexpCnvBoundaries(gexp_norm, genes: GRanges, m=0.15, chrs=None, min_traverse=3, t=1e-6,
                          min_num_genes=3, trim=0.1, verbose=False, grouping=None):
    """R/HoneyBADGER_gexp.R:462-584.  Returns dict(amp=dict(info, regions) | None, del=...)."""
    chrs = CHRS_DEFAULT if chrs is None else list(chrs)
    df = _as_frame(gexp_norm)
    ordr = order_genes(list(df.index), genes, chrs, has_na=df.isna().any(axis=1).values)
    df = df.iloc[ordr]
    mat = df.values
    names = np.asarray(df.index, dtype=object)
    G, N = mat.shape
    heights = list(range(1, min(min_traverse, N) + 1))
    x_dev = _to_dev_f64(mat)
    labels = grouping(mat, heights) if grouping else ward_groups_dev(x_dev, heights, 101)
    member, _ = _membership(labels, N)
    votes = {"amp": np.zeros(G, np.int64), "del": np.zeros(G, np.int64)}
    if member.shape[0]:
        y, _, _ = gexp_round_dev(x_dev, member, [-m, 0.0, m], t)
        votes["amp"] = (y == 3).sum(0)
        votes["del"] = (y == 1).sum(0)

    def tbv(vote):
        if verbose:
            print(f"max vote:{vote.max()}")
        runs = runs_from_vote(vote, min_num_genes, trim)
        if runs is None:
            return None
        info = [list(names[i]) for i in runs]
        regions = [r for bs in info for r in genes[bs].range()]
        return {"info": info, "regions": regions}

    return {"amp": tbv(votes["amp"]), "del": tbv(votes["del"])}


def calcAlleleCnvBoundaries(r_maf, n_sc, l_maf=None, n_bulk=None, snps: GRanges | None = None,
                            geneFactor=None, min_traverse=3, t=1e-6, pd=0.1, pn=0.45,
                            min_num_snps=5, trim=0.1, verbose=False, grouping=None):
    """R/HoneyBADGER_allele.R:543-651.  Only the first group of each height votes (:594-596)."""
    rdf, ndf = _as_frame(r_maf, "snp"), _as_frame(n_sc, "snp")
    names = np.asarray(rdf.index, dtype=object)
    snps = snps if snps is not None else snps_from_names(names)
    r, n = rdf.values, ndf.values
    G, N = r.shape
    heights = list(range(1, min(min_traverse, N) + 1))
    r_dev, n_dev = _to_dev_i32(r), _to_dev_i32(n)
    if grouping is None:
        labels = ward_groups_dev(None, heights, 31, counts=(r_dev, n_dev))
    else:
        labels = grouping(r, n, heights)
    member, tags = _membership(labels, N)
    vote = np.zeros(G, np.int64)
    if member.shape[0]:
        y, _, _ = allele_round_dev(r_dev, n_dev, member, pd, pn, t)
        for k, (_, gi) in enumerate(tags):
            if gi == 0:
                vote += y[k] == 1
    if verbose:
        print(f"max vote:{vote.max()}")
    runs = runs_from_vote(vote, min_num_snps, trim)
    if runs is None:
        return None
    info = [list(names[i]) for i in runs]
    return {"info": info, "regions": [rg for bs in info for rg in snps[bs].range()]}


# --------------------------------------------------------------------------- RefClass mirror
class HoneyBADGER:
    """setRefClass("HoneyBADGER", ...) (R/HoneyBADGER.R:30-94): the fields and methods on the
    segmentation path.  Matrices live on the GPU after set*Mats; recursion passes index sets."""

    def __init__(self, name=None):
        self.name = name
        self.gexp_norm = None
        self.genes = None
        self.dev = None
        self.mvFit = None
        self.r_maf = self.n_sc = self.l_maf = self.n_bulk = self.snps = None
        self.pred_genes_r = self.pred_snps_r = None
        self.bound_genes_old = None
        self.bound_genes_final = None
        self.bound_snps_old = None
        self.bound_snps_final = None
        self.grouping_gexp = None   # None: ward_groups_dev on the resident matrix; or f(mat, heights)
        self.grouping_allele = None
        self.retest_gexp = retest_gexp_closed_form
        self.retest_allele = None   # None: retest_allele_model (snpModel.bug marginal); or f(r, n, pd, pn)
        self.geneFactor = None      # SNP name -> gene (setGeneFactors needs TxDb; pass a dict); None: 1 SNP = 1 gene
        self._x_dev = self._r_dev = self._n_dev = None
        self._x_any_nan = None

    # ---- inputs (kept as the reference defines them; R/HoneyBADGER.R:123-188, :485-614)
    def setGexpMats(self, gexp_sc_init, gexp_ref_init, genes: GRanges, filter=True, minMeanBoth=0,
                    minMeanTest=None, minMeanRef=None, scale=True, verbose=True):
        """R/HoneyBADGER.R:124-188.  `genes` replaces the biomaRt lookup (mart.obj needs the network);
        minMeanTest / minMeanRef None = the reference's defaults mean(m[m != 0]) of the init matrices.
        The arithmetic runs on the device (`set_gexp_mats_dev`); the result stays resident."""
        import torch
        if verbose:
            print("Initializing expression matrices ... ")
        sc = _as_frame(gexp_sc_init)
        ref = gexp_ref_init if isinstance(gexp_ref_init, pd.DataFrame) else \
            pd.DataFrame(np.asarray(gexp_ref_init, dtype=np.float64).reshape(len(sc.index), -1),
                         index=getattr(gexp_ref_init, "index", sc.index))
        inref = set(ref.index)
        vi = [g for g in sc.index if g in inref]
        if len(vi) < 10:
            print("WARNING! GENE NAMES IN EXPRESSION MATRICES DO NOT SEEM TO MATCH! ")
        if filter and (len(vi) != len(sc.index) or len(vi) != len(ref.index)):
            # the default thresholds are promises on the *init* matrices: when the name intersection
            # drops rows they are not the library's defaults (taken on the matrices it is handed), so
            # evaluate them here with R's serial mean replayed exactly
            lib = _lib.load()
            for which, m in (("t", sc), ("r", ref)):
                if (minMeanTest if which == "t" else minMeanRef) is not None:
                    continue
                d = _to_dev_f64(m.values)
                mean = C.c_double(0)
                _lib.check(lib.hb_mean_nonzero_dev(d.data_ptr(), d.stride(0), d.shape[0], d.shape[1], 1,
                                                   C.byref(mean), None, _stream()))
                if which == "t":
                    minMeanTest = mean.value
                else:
                    minMeanRef = mean.value
        sc, ref = sc.loc[vi], ref.loc[vi]
        norm_dev, keep, _ = set_gexp_mats_dev(sc.values, ref.values, filter, minMeanBoth, minMeanTest,
                                              minMeanRef, scale)
        if filter and verbose:
            print(f"{keep.size} genes passed filtering ... ")
        if scale and verbose:
            print("Scaling coverage ... ")
        if verbose:
            print(f"Normalizing gene expression for {keep.size} genes and {sc.shape[1]} cells ... ")
        names = np.asarray(sc.index, dtype=object)[keep]
        has_pos = np.array([g in genes._ix for g in names], dtype=bool)
        gvi = list(names[has_pos])
        self.genes = genes[gvi]
        if not has_pos.all():
            norm_dev = norm_dev.index_select(0, torch.as_tensor(np.nonzero(has_pos)[0], device=norm_dev.device))
        norm_dev = norm_dev.contiguous()
        self.gexp_norm = pd.DataFrame(norm_dev.cpu().numpy(), index=gvi, columns=sc.columns)
        self._x_dev = (norm_dev, {g: i for i, g in enumerate(gvi)}, {c: i for i, c in enumerate(sc.columns)})
        self._x_any_nan = None
        if verbose:
            print("Done setting initial expression matrices! ")

    def setGexpDev(self, dev=None, alpha=0.05, n=100, seed=0, verbose=False):
        """R/HoneyBADGER.R:1054-1077 draws from R's RNG (ks.test of rnorm samples); the value is an
        input here (the vignette's is 1.015658, docs/Getting_Started.md:108)."""
        if dev is None:
            raise NotImplementedError("setGexpDev needs R's RNG stream; pass dev= explicitly")
        self.dev = float(dev)
        if verbose:
            print(f"Optimal deviance: {self.dev}")

    def setAlleleMats(self, r_init, n_sc_init, l_init=None, n_bulk_init=None, filter=True,
                      het_deviance_threshold=0.05, min_cell=3, verbose=True):
        """R/HoneyBADGER.R:486-614; counting, filtering and the lesser-allele flip run on the device
        (`set_allele_mats_dev`) and the matrices stay resident."""
        if verbose:
            print("Initializing allele matrices ... ")
        r, n = _as_frame(r_init, "snp"), _as_frame(n_sc_init, "snp")
        if verbose and (l_init is None or n_bulk_init is None):
            print("Creating in-silico bulk ... ")
            print(f"using {r.shape[1]} cells ... ")
        out = set_allele_mats_dev(r.values, n.values, l_init, n_bulk_init, filter, het_deviance_threshold,
                                  min_cell)
        keep = out["keep"]
        if filter:
            if verbose:
                print("Filtering for putative heterozygous snps ... ")
                print(f"allowing for a {het_deviance_threshold} deviation from the expected 0.5 heterozygous allele fraction ... ")
                print(f"must have coverage in at least {min_cell} cells ... ")
            # the reference prints length(vi) of the coverage filter = SNPs that passed the first one
            print(f"{out['n_het']} heterozygous SNPs identified ")
        names = list(np.asarray(r.index, dtype=object)[keep])
        self.r_maf = pd.DataFrame(out["r_maf"].cpu().numpy(), index=names, columns=r.columns)
        self.n_sc = pd.DataFrame(out["n_sc"].cpu().numpy(), index=names, columns=r.columns)
        self.l_maf = out["l_maf"]
        self.n_bulk = out["n_bulk"]
        self.snps = snps_from_names(names)
        self._r_dev = (out["r_maf"].contiguous(), out["n_sc"].contiguous(), {g: i for i, g in enumerate(names)},
                       {c: i for i, c in enumerate(r.columns)})
        if verbose:
            print("Done setting initial allele matrices! ")

    # ---- sub-matrices of the recursion: names only, the values stay on the device
    class _View:
        """`m[rows, cols]` of a resident matrix by name — what the recursive calls hand down instead
        of a host copy (R/HoneyBADGER.R:1304-1309 passes `gexp.norm[, g1]`)."""

        def __init__(self, index, columns):
            self.index = np.asarray(index, dtype=object)
            self.columns = np.asarray(columns, dtype=object)

    # ---- device residency: full matrices are uploaded once; every level gathers rows / columns there
    def _resident(self, kind):
        import torch
        if kind == "gexp":
            if self._x_dev is None:
                self._x_dev = (_to_dev_f64(self.gexp_norm.values),
                               {g: i for i, g in enumerate(self.gexp_norm.index)},
                               {c: i for i, c in enumerate(self.gexp_norm.columns)})
            return self._x_dev
        if self._r_dev is None:
            self._r_dev = (_to_dev_i32(self.r_maf.values), _to_dev_i32(self.n_sc.values),
                           {g: i for i, g in enumerate(self.r_maf.index)},
                           {c: i for i, c in enumerate(self.r_maf.columns)})
        return self._r_dev

    @staticmethod
    def _gather(full, rix, cix, rows, cols):
        """full[rows, cols] on the device, or None when a name is not resident (user sub-matrix)."""
        import torch
        try:
            ri = torch.as_tensor([rix[r] for r in rows], device=full.device)
            ci = torch.as_tensor([cix[c] for c in cols], device=full.device)
        except KeyError:
            return None
        return full.index_select(0, ri).index_select(1, ci).contiguous()

    # ---- expression boundaries (R/HoneyBADGER.R:1091-1316)
    def calcGexpCnvBoundaries(self, gexp_norm_sub=None, chrs=None, min_traverse=5, min_num_genes=10,
                              t=1e-6, init=False, verbose=False, **retest_kw):
        import torch
        chrs = CHRS_DEFAULT if chrs is None else list(chrs)
        sub = self.gexp_norm if gexp_norm_sub is None else gexp_norm_sub
        rows_all = np.asarray(sub.index, dtype=object)
        cols = np.asarray(sub.columns, dtype=object)
        if init:
            self.pred_genes_r = _zeros_frame(rows_all, cols)
            self.bound_genes_old = []
            self.bound_genes_final = []
        if self.bound_genes_final is None:
            print("ERROR! USE init=TRUE! ")
            return None
        old = set(self.bound_genes_old)
        keep = np.fromiter((g not in old for g in rows_all), dtype=bool, count=rows_all.size)
        rows = rows_all[keep]
        # the values never leave the device: rows / columns are gathered by name from the resident
        # matrix; only a sub-matrix with names that are not resident is uploaded
        x_dev = None
        resident = self.gexp_norm is not None
        if resident:
            full, rix, cix = self._resident("gexp")
            if self._x_any_nan is None:
                self._x_any_nan = bool(torch.isnan(full).any().item())
            if not self._x_any_nan:  # na.omit has nothing to drop: order by name, gather once
                ordr = order_genes(list(rows), self.genes, chrs)
                x_dev = self._gather(full, rix, cix, rows[ordr], cols)
            else:
                x_dev = self._gather(full, rix, cix, rows, cols)
        fresh = x_dev is None
        if fresh:
            if isinstance(sub, HoneyBADGER._View):
                raise KeyError("sub-matrix names are not resident")
            x_dev = _to_dev_f64(sub.values[keep])
        if fresh or self._x_any_nan:
            has_na = torch.isnan(x_dev).any(dim=1).cpu().numpy()
            ordr = order_genes(list(rows), self.genes, chrs, has_na=has_na)
            x_dev = x_dev.index_select(0, torch.as_tensor(ordr, device=x_dev.device))
        names = rows[ordr]
        G, N = x_dev.shape
        heights = list(range(1, min(min_traverse, N) + 1))
        labels = self.grouping_gexp(x_dev.cpu().numpy(), heights) if self.grouping_gexp else \
            ward_groups_dev(x_dev, heights, 101)
        member, _ = _membership(labels, N)
        amp_v, del_v = np.zeros(G, np.int64), np.zeros(G, np.int64)
        if member.shape[0]:
            y, _, _ = gexp_round_dev(x_dev, member, [-self.dev, 0.0, self.dev], t)
            amp_v, del_v = (y == 3).sum(0), (y == 1).sum(0)

        def tbv(vote, which):
            # `vote[bound.genes.old] <- 0` only appends zero entries: the old rows are gone already
            if verbose:
                print(f"max vote:{vote.max()}")
            runs = runs_from_vote(vote, min_num_genes, None)  # the method does not trim
            if runs is None:
                return []
            out = []
            for idx in runs:
                ap, dp = self.retest_gexp(x_dev[idx], self.dev, **retest_kw)
                out.append({"ap": ap, "dp": dp, "bs": list(names[idx])})
            return out

        amp_info, del_info = tbv(amp_v, "amp"), tbv(del_v, "del")
        rows = [d["dp"] for d in del_info] + [a["ap"] for a in amp_info]
        lists = [d["bs"] for d in del_info] + [a["bs"] for a in amp_info]
        if not rows:
            return None
        prob = np.vstack(rows)
        dps = (prob > 0.75).sum(1)  # rowSums(prob.bin, na.rm=TRUE): only the 1s count
        dpsi = int(np.nonzero(dps == dps.max())[0][0])
        prob_fin, bound_new = prob[dpsi], lists[dpsi]
        cells = cols
        g1, g2 = list(cells[prob_fin > 0.75]), list(cells[prob_fin <= 0.25])
        if g1:
            self.pred_genes_r.loc[bound_new, g1] = 1
            self.bound_genes_final.append([bound_new])
        self.bound_genes_old = list(dict.fromkeys(bound_new + self.bound_genes_old))
        for g in (g1, g2):  # the recursive calls use the DEFAULT arguments (R/HoneyBADGER.R:1304)
            if len(g) >= 3:
                try:
                    self.calcGexpCnvBoundaries(gexp_norm_sub=HoneyBADGER._View(names, g) if resident and not fresh
                                               else sub.loc[list(names), g])
                except Exception as e:  # tryCatch(..., error = function(e) cat(...))
                    print(f"ERROR: {e}")
        return None

    # ---- allele boundaries (R/HoneyBADGER.R:1336-1578)
    def calcAlleleCnvBoundaries(self, r_sub=None, n_sc_sub=None, l_sub=None, n_bulk_sub=None,
                                min_traverse=3, t=1e-6, pd=0.1, pn=0.45, min_num_snps=5, trim=0.1,
                                init=False, verbose=False, **retest_kw):
        r_maf = self.r_maf if r_sub is None else r_sub
        n_sc = self.n_sc if n_sc_sub is None else n_sc_sub
        rows_all = np.asarray(r_maf.index, dtype=object)
        cols = np.asarray(r_maf.columns, dtype=object)
        if init:
            self.pred_snps_r = _zeros_frame(rows_all, cols)  # (`pd` is the deletion probability in here)
            self.bound_snps_old = []
            self.bound_snps_final = []
        if self.bound_snps_final is None:
            print("ERROR! USE init=TRUE! ")
            return None
        old = set(self.bound_snps_old)
        vi = np.fromiter((s not in old for s in rows_all), dtype=bool, count=rows_all.size)
        names = rows_all[vi]
        r_dev = n_dev = None
        resident = self.r_maf is not None
        if resident:  # gather by name on the device; the recursion hands down names, not copies
            rfull, nfull, rix, cix = self._resident("allele")
            r_dev = self._gather(rfull, rix, cix, names, cols)
            n_dev = self._gather(nfull, rix, cix, names, cols)
        fresh = r_dev is None
        if fresh:
            if isinstance(r_maf, HoneyBADGER._View):
                raise KeyError("sub-matrix names are not resident")
            r_dev, n_dev = _to_dev_i32(r_maf.values[vi]), _to_dev_i32(n_sc.values[vi])
        G, N = r_dev.shape
        heights = list(range(1, min(min_traverse, N) + 1))
        if self.grouping_allele is None:
            labels = ward_groups_dev(None, heights, 31, counts=(r_dev, n_dev))
        else:
            labels = self.grouping_allele(r_dev.cpu().numpy(), n_dev.cpu().numpy(), heights)
        member, tags = _membership(labels, N)
        vote = np.zeros(G, np.int64)
        if member.shape[0]:
            y, _, _ = allele_round_dev(r_dev, n_dev, member, pd, pn, t)
            for k, (_, gi) in enumerate(tags):
                if gi == 0:  # vote[b[[1]]]: only the first group of each height (:1424-1426)
                    vote += y[k] == 1
        if verbose:
            print(f"max vote:{vote.max()}")
        runs = runs_from_vote(vote, min_num_snps, trim)
        if runs is None:
            return None
        def bulk_for(idx):
            # l.maf / n.bulk of the region (R/HoneyBADGER.R:1486): the object's fields at the top level,
            # l.sub / n.bulk.sub when given, else what the recursion passes down (:1563-1566):
            # rowSums(r.maf[, g] > 0), rowSums(n.sc[, g] > 0) — counted on the device
            if r_sub is None and self.l_maf is not None:
                ii = [self._r_dev[2][s] for s in names[idx]]
                return np.asarray(self.l_maf, dtype=np.float64)[ii], np.asarray(self.n_bulk, dtype=np.float64)[ii]
            if l_sub is not None and n_bulk_sub is not None:
                return (np.asarray(l_sub, dtype=np.float64)[vi][idx],
                        np.asarray(n_bulk_sub, dtype=np.float64)[vi][idx])
            return ((r_dev[idx] > 0).sum(1).double().cpu().numpy(), (n_dev[idx] > 0).sum(1).double().cpu().numpy())

        def retest(idx):
            if self.retest_allele is not None:  # user hook f(r_rows, n_rows, pd, pn, ...)
                return self.retest_allele(r_dev[idx], n_dev[idx], pd, pn, **retest_kw)
            lb, nbk = bulk_for(idx)
            gf = None if self.geneFactor is None else [self.geneFactor.get(s, s) for s in names[idx]]
            return retest_allele_model(r_dev[idx], n_dev[idx], lb, nbk, gene=gf, pe=pd, **retest_kw)

        probs = [retest(idx) for idx in runs]
        lists = [list(names[idx]) for idx in runs]
        prob = np.vstack(probs)
        if len(runs) > 1:
            dps = (prob > 0.75).sum(1)
            dpsi = int(np.nonzero(dps == dps.max())[0][0])
            prob_fin, bound_new = prob[dpsi], lists[dpsi]
        else:
            prob_fin, bound_new = prob[0], lists[0]
        cells = cols
        g1, g2 = list(cells[prob_fin > 0.75]), list(cells[prob_fin <= 0.25])
        if g1:
            self.pred_snps_r.loc[bound_new, g1] = 1
            self.bound_snps_final.append([bound_new])
        self.bound_snps_old = list(dict.fromkeys(bound_new + self.bound_snps_old))
        for g in (g1, g2):
            if len(g) >= 3:
                try:
                    # l.sub / n.bulk.sub (R/HoneyBADGER.R:1563-1566) feed only the bulk term of the retest (:1486):
                    # the next level counts them on the device from the same sub-matrices (bulk_for)
                    if resident and not fresh:
                        v = HoneyBADGER._View(names, g)
                        self.calcAlleleCnvBoundaries(r_sub=v, n_sc_sub=v)
                    else:
                        self.calcAlleleCnvBoundaries(r_sub=r_maf.loc[list(names), g], n_sc_sub=n_sc.loc[list(names), g])
                except Exception as e:
                    print(f"ERROR: {e}")
        return None


def _zeros_frame(index, columns) -> pd.DataFrame:
    """matrix(0, nrow, ncol) with dimnames; calloc-backed, not copied."""
    return pd.DataFrame(np.zeros((len(index), len(columns))), index=index, columns=columns, copy=False)
