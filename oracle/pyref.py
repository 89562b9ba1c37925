"""Independent pure-Python restatement of the HMM segmentation path — TEST INFRASTRUCTURE ONLY.

A second, deliberately naive implementation (Python floats are IEEE doubles, one rounding per
operation, no FMA) used to cross-check oracle/hb_oracle.c, to generate tests/golden, and to hold
the literal restatement of the reference's driver loops (vote / runs / filter / trim), which are
read directly from the reference:

  R/HoneyBADGER_gexp.R:462-584      calcGexpCnvBoundaries   (function, deterministic)
  R/HoneyBADGER_allele.R:543-651    calcAlleleCnvBoundaries (function, deterministic)
  R/HoneyBADGER.R:1091-1316, :1336-1578   RefClass methods (recursive; retest is JAGS)

PARITY UNPINNED for the HiddenMarkov / nmath parts (see hb_oracle.c).  mpmath versions of the
densities give an arithmetic-independent check of their values.
"""
from __future__ import annotations

import math

import numpy as np

M_LN_SQRT_2PI = 0.918938533204672741780329736406
M_LN_2PI = 1.837877066409345483560659472811
DBL_MIN = 2.2250738585072014e-308

SFERR_HALVES = [
    0.0, 0.1534264097200273452913848, 0.0810614667953272582196702,
    0.0548141210519176538961390, 0.0413406959554092940938221, 0.03316287351993628748511048,
    0.02767792568499833914878929, 0.02374616365629749597132920, 0.02079067210376509311152277,
    0.01848845053267318523077934, 0.01664469118982119216319487, 0.01513497322191737887351255,
    0.01387612882307074799874573, 0.01281046524292022692424986, 0.01189670994589177009505572,
    0.01110455975820691732662991, 0.010411265261972096497478567, 0.009799416126158803298389475,
    0.009255462182712732917728637, 0.008768700134139385462952823, 0.008330563433362871256469318,
    0.007934114564314020547248100, 0.007573675487951840794972024, 0.007244554301320383179543912,
    0.006942840107209529865664152, 0.006665247032707682442354394, 0.006408994188004207068439631,
    0.006171712263039457647532867, 0.005951370112758847735624416, 0.005746216513010115682023589,
    0.005554733551962801371038690,
]


# --------------------------------------------------------------------------- densities
def dnorm_log(x: float, mu: float, sigma: float) -> float:
    z = abs((x - mu) / sigma)
    return -(M_LN_SQRT_2PI + 0.5 * z * z + math.log(sigma))


def stirlerr(n: float) -> float:
    S0, S1, S2, S3, S4 = (0.083333333333333333333, 0.00277777777777777777778,
                          0.00079365079365079365079365, 0.000595238095238095238095238,
                          0.0008417508417508417508417508)
    if n <= 15.0:
        nn = n + n
        if nn == int(nn):
            return SFERR_HALVES[int(nn)]
        return math.lgamma(n + 1.0) - (n + 0.5) * math.log(n) + n - M_LN_SQRT_2PI
    nn = n * n
    if n > 500:
        return (S0 - S1 / nn) / n
    if n > 80:
        return (S0 - (S1 - S2 / nn) / nn) / n
    if n > 35:
        return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n


def bd0(x: float, np_: float) -> float:
    if abs(x - np_) < 0.1 * (x + np_):
        v = (x - np_) / (x + np_)
        s = (x - np_) * v
        if abs(s) < DBL_MIN:
            return s
        ej = 2 * x * v
        v = v * v
        for j in range(1, 1000):
            ej *= v
            s1 = s + ej / ((j << 1) + 1)
            if s1 == s:
                return s1
            s = s1
    return x * math.log(x / np_) + np_ - x


def dbinom_log(x: float, n: float, p: float) -> float:
    q = 1 - p
    if p == 0:
        return 0.0 if x == 0 else -math.inf
    if q == 0:
        return 0.0 if x == n else -math.inf
    if x == 0:
        if n == 0:
            return 0.0
        return (-bd0(n, n * q) - n * p) if p < 0.1 else n * math.log(q)
    if x == n:
        return (-bd0(n, n * p) - n * q) if q < 0.1 else n * math.log(p)
    if x < 0 or x > n:
        return -math.inf
    lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q)
    lf = M_LN_2PI + math.log(x) + math.log1p(-x / n)
    return lc - 0.5 * lf


def mp_dnorm_log(x, mu, sigma, dps=50):
    import mpmath as mp
    with mp.workdps(dps):
        z = (mp.mpf(x) - mp.mpf(mu)) / mp.mpf(sigma)
        return -(mp.log(mp.sqrt(2 * mp.pi)) + z * z / 2 + mp.log(mp.mpf(sigma)))


def mp_dbinom_log(x, n, p, dps=50):
    import mpmath as mp
    with mp.workdps(dps):
        x, n = int(x), int(n)
        if x < 0 or x > n:
            return -mp.inf
        P = mp.mpf(p)
        return (mp.log(mp.binomial(n, x)) + (x * mp.log(P) if x else 0)
                + ((n - x) * mp.log(1 - P) if n - x else 0))


def betabinom_logpmf_exact(x, n, p, rho, dps=60):
    """Exact (mpmath) log beta-binomial pmf with mean p and intra-class correlation rho:
    log C(n,x) + log B(x + a, n - x + b) - log B(a, b), a = p(1-rho)/rho, b = (1-p)(1-rho)/rho — an independent
    formulation (log-gamma at high precision) of what hb_oracle.c's hbo_dbetabinom_log evaluates as
    binomial + correction."""
    import mpmath as mp
    with mp.workdps(dps):
        x, n = int(x), int(n)
        if x < 0 or x > n:
            return -mp.inf
        P, R = mp.mpf(p), mp.mpf(rho)
        if R == 0:
            return mp_dbinom_log(x, n, p, dps)
        s = (1 - R) / R
        a, b = P * s, (1 - P) * s
        lbeta = lambda u, v: mp.loggamma(u) + mp.loggamma(v) - mp.loggamma(u + v)  # noqa: E731
        return mp.log(mp.binomial(n, x)) + lbeta(x + a, n - x + b) - lbeta(a, b)


# --------------------------------------------------------------------------- base-R reductions
def r_mean(x) -> float:
    """summary.c real_mean: x87 LDOUBLE, two passes, sequential order."""
    xl = np.asarray(x, dtype=np.float64).astype(np.longdouble)
    n = np.longdouble(xl.size)
    s = np.cumsum(xl)[-1] / n
    t = np.cumsum(xl - s)[-1]
    return float(s + t / n)


def r_sd(x) -> float:
    """sqrt(var(x)); cov.c cov_complete1: mean rounded to double, LDOUBLE sum of squares, n-1."""
    x = np.asarray(x, dtype=np.float64)
    xm = np.longdouble(r_mean(x))
    d = x.astype(np.longdouble) - xm
    ss = np.cumsum(d * d)[-1]
    return math.sqrt(float(ss / np.longdouble(x.size - 1)))


# --------------------------------------------------------------------------- HiddenMarkov
def viterbi_logb(logb, Pi, delta):
    """Literal HiddenMarkov::Viterbi.dthmm on a log-density table logb[T][S]; 1-based states."""
    logb = np.asarray(logb, dtype=np.float64)
    T, S = logb.shape
    with np.errstate(divide="ignore"):
        logPi = [[math.log(v) if v > 0 else -math.inf for v in row] for row in np.asarray(Pi)]
        logdelta = [math.log(v) if v > 0 else -math.inf for v in delta]
    nu = [[0.0] * S for _ in range(T)]
    for k in range(S):
        nu[0][k] = logdelta[k] + float(logb[0][k])
    for i in range(1, T):
        prev = nu[i - 1]
        for k in range(S):
            nu[i][k] = max(prev[j] + logPi[j][k] for j in range(S)) + float(logb[i][k])
    if any(v == -math.inf for v in nu[T - 1]):
        raise FloatingPointError("Problems With Underflow")
    y = [0] * T

    def which_max(v):
        b = 0
        for j in range(1, len(v)):
            if v[j] > v[b]:
                b = j
        return b

    y[T - 1] = which_max(nu[T - 1])
    for i in range(T - 2, -1, -1):
        y[i] = which_max([logPi[j][y[i + 1]] + nu[i][j] for j in range(S)])
    return np.asarray(y, dtype=np.int32) + 1


def viterbi_norm(x, mean, sd, Pi, delta):
    logb = [[dnorm_log(float(xi), float(mean[k]), float(sd[k])) for k in range(len(mean))]
            for xi in x]
    return viterbi_logb(logb, Pi, delta)


def viterbi_binom(x, size, prob, Pi, delta):
    logb = [[dbinom_log(float(xi), float(ni), float(p)) for p in prob] for xi, ni in zip(x, size)]
    return viterbi_logb(logb, Pi, delta)


def forwardback_prob(prob, Pi, delta):
    """Literal HiddenMarkov::forwardback.dthmm (scaled linear domain) + Estep's u; numpy fp64."""
    prob = np.asarray(prob, dtype=np.float64)
    Pi = np.asarray(Pi, dtype=np.float64)
    T, S = prob.shape
    la = np.zeros((T, S))
    lb = np.zeros((T, S))
    phi = np.asarray(delta, dtype=np.float64).copy()
    lscale = 0.0
    with np.errstate(divide="ignore"):
        for i in range(T):
            if i > 0:
                phi = phi @ Pi
            phi = phi * prob[i]
            s = phi.sum()
            phi = phi / s
            lscale += math.log(s)
            la[i] = np.log(phi) + lscale
        LL = lscale
        phi = np.full(S, 1.0 / S)
        lscale = math.log(S)
        for i in range(T - 2, -1, -1):
            phi = Pi @ (prob[i + 1] * phi)
            lb[i] = np.log(phi) + lscale
            s = phi.sum()
            phi = phi / s
            lscale += math.log(s)
    return np.exp(la + lb - LL), LL


def mp_posteriors(logb, Pi, delta, dps=60):
    """Exact-arithmetic (mpmath) state posteriors from a log-density table; small T only."""
    import mpmath as mp
    with mp.workdps(dps):
        T, S = len(logb), len(logb[0])
        b = [[mp.e ** mp.mpf(float(v)) if v > -math.inf else mp.mpf(0) for v in row]
             for row in logb]
        P = [[mp.mpf(float(v)) for v in row] for row in np.asarray(Pi)]
        a = [[mp.mpf(0)] * S for _ in range(T)]
        for k in range(S):
            a[0][k] = mp.mpf(float(delta[k])) * b[0][k]
        for i in range(1, T):
            for k in range(S):
                a[i][k] = sum(a[i - 1][j] * P[j][k] for j in range(S)) * b[i][k]
            tot = sum(a[i])
            a[i] = [v / tot for v in a[i]]
        be = [[mp.mpf(1)] * S for _ in range(T)]
        for i in range(T - 2, -1, -1):
            for j in range(S):
                be[i][j] = sum(P[j][k] * b[i + 1][k] * be[i + 1][k] for k in range(S))
            tot = sum(be[i])
            be[i] = [v / tot for v in be[i]]
        g = np.zeros((T, S))
        for i in range(T):
            w = [a[i][k] * be[i][k] for k in range(S)]
            tot = sum(w)
            g[i] = [float(v / tot) for v in w]
        return g


# --------------------------------------------------------------------------- driver loops
def r_round_half_even(v: float) -> int:
    return int(round(v))  # Python's round() is IEC 60559 half-even, like R's for digits=0


def runs_filter_trim(vote, min_num: int, trim: float | None):
    """vote -> list of 0-based index arrays (one per surviving run), or None.

    Literal restatement of R/HoneyBADGER.R:1178-1210 (+ trim :1476 / _gexp.R:568):
      * `max(vote)==0` -> NULL
      * binarise; label runs with the `for(i in 2:length(vote))` loop — the first element of every
        run keeps label 0
      * table() without label 0; drop labels with count < min_num; none left -> NULL
      * per label: names[cont==label], then names[1:round(L - L*trim)] when trim is not None
    """
    vote = np.asarray(vote)
    n = vote.size
    if n == 0 or vote.max() == 0:
        return None
    v = (vote > 0).astype(np.int64)
    cs = 1
    cont = np.zeros(n, dtype=np.int64)
    for i in range(1, n):
        if v[i] >= 1 and v[i] == v[i - 1]:
            cont[i] = cs
        else:
            cs += 1
    labels, counts = np.unique(cont, return_counts=True)  # sorted like table()
    labels, counts = labels[1:], counts[1:]  # `tbv[-1]`: drop the first level (label 0)
    keep = counts >= min_num
    if not keep.any():
        return None
    out = []
    for lab in labels[keep]:
        idx = np.nonzero(cont == lab)[0]
        if trim is not None:
            L = idx.size
            r = r_round_half_even(L - L * trim)
            if r >= 1:
                idx = idx[:r]
            elif r == 0:
                idx = idx[:1]  # R: x[1:0] == x[c(1,0)] == x[1]
            else:
                raise IndexError("only 0's may be mixed with negative subscripts")
        out.append(idx)
    return out


def group_members(labels):
    """unique(ct) order = first appearance; returns list of (group id, 0-based cell index array)."""
    labels = np.asarray(labels)
    seen = []
    for v in labels:
        if v not in seen:
            seen.append(v)
    return [(g, np.nonzero(labels == g)[0]) for g in seen]


def gexp_hmm_calls(gexp_norm, labels_per_height, dev, t=1e-6):
    """The `lapply(heights, ...)` block (R/HoneyBADGER.R:1135-1162): one dict(amp, del) of 0-based
    gene indices per (height, group with >1 cell), flattened in (height, group) order."""
    from . import oracle as O
    gexp_norm = np.asarray(gexp_norm, dtype=np.float64)
    Pi = O.sym_Pi(3, t)
    out = []
    for labels in labels_per_height:
        for _, cells in group_members(labels):
            if cells.size > 1:
                x = O.pool_mean(gexp_norm, cells)
                sd = O.r_sd(x)
                y = O.viterbi_norm(x, [-dev, 0.0, dev], [sd, sd, sd], Pi, [0.0, 1.0, 0.0])
                out.append({"amp": np.nonzero(y == 3)[0], "del": np.nonzero(y == 1)[0],
                            "x": x, "sd": sd, "y": y})
    return out


def calc_gexp_cnv_boundaries_fn(gexp_norm, labels_per_height, m=0.15, t=1e-6, min_num_genes=3,
                                trim=0.1):
    """R/HoneyBADGER_gexp.R:462-584 after ordering, with hclust labels supplied.  Returns
    dict(amp=[idx arrays] | None, del=[...] | None) of 0-based row indices."""
    G = np.asarray(gexp_norm).shape[0]
    calls = gexp_hmm_calls(gexp_norm, labels_per_height, m, t)

    def get_tbv(key):
        vote = np.zeros(G, dtype=np.int64)
        for c in calls:
            vote[c[key]] += 1
        return runs_filter_trim(vote, min_num_genes, trim)

    return {"amp": get_tbv("amp"), "del": get_tbv("del"), "calls": calls}


def allele_hmm_calls(r_maf, n_sc, labels_per_height, t=1e-6, pd=0.1, pn=0.45):
    """`lapply(heights, ...)` of R/HoneyBADGER.R:1395-1417: list over heights of lists over groups
    (None for single-cell groups) of 0-based SNP indices in state 1."""
    from . import oracle as O
    Pi = O.sym_Pi(2, t)
    out = []
    for labels in labels_per_height:
        per_group = []
        for _, cells in group_members(labels):
            if cells.size > 1:
                mafl = O.pool_count_pos(r_maf, cells)
                sizel = O.pool_count_pos(n_sc, cells)
                y = O.viterbi_binom(mafl, sizel, [pd, pn], Pi, [0.0, 1.0])
                per_group.append({"snps": np.nonzero(y == 1)[0], "mafl": mafl, "sizel": sizel,
                                  "y": y})
            else:
                per_group.append(None)
        out.append(per_group)
    return out


def calc_allele_cnv_boundaries_fn(r_maf, n_sc, labels_per_height, t=1e-6, pd=0.1, pn=0.45,
                                  min_num_snps=5, trim=0.1):
    """R/HoneyBADGER_allele.R:543-651 with hclust labels supplied; 0-based SNP index arrays.
    Only the first group of each height votes (`vote[b[[1]]]`, :594-596)."""
    G = np.asarray(r_maf).shape[0]
    calls = allele_hmm_calls(r_maf, n_sc, labels_per_height, t, pd, pn)
    vote = np.zeros(G, dtype=np.int64)
    for per_group in calls:
        if per_group and per_group[0] is not None:
            vote[per_group[0]["snps"]] += 1
    return {"info": runs_filter_trim(vote, min_num_snps, trim), "calls": calls}
