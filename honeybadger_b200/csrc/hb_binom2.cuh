// hb_binom2.cuh — 2-state binomial chain (deletion / neutral): fused Viterbi + forward-backward,
// one thread per chain.
//
// Replaces HiddenMarkov::dthmm(mafl, Pi, delta, "binom", list(prob=c(pd,pn)), list(size=sizel),
// discrete=TRUE) -> HiddenMarkov::Viterbi at R/HoneyBADGER.R:1409-1410 and
// R/HoneyBADGER_allele.R:579-580.
//
// Emissions: log dbinom(x, n, p_k) comes from a table indexed by (n, x) for n <= lut_nmax, built
// by the library's HOST code with the host libm, so those values are bit-identical to what R's
// nmath produces on the same machine (zero-coverage SNPs give (0, 0): the source of the
// structural ties the reference resolves by "first index wins").  Counts above the table are
// evaluated in-kernel with the same saddle-point algorithm (device log/log1p: values may differ
// from the host's in the last ulp).
// Same two-sweep structure as hb_norm3.cuh; slots per step for the backward sweep: phi[2], rho.
#pragma once
#include "hb_device.cuh"
#include "hb_norm3.cuh" // HB_CHUNK

namespace hb {

#define HB_BINOM2_SLOT 4 // per step for the backward sweep: m0, m1, (d | er), mr

// Emission ratio b_0/b_1 = m * 2^e with m in [0.7, 1.42).
struct RhoME {
    double m;
    int32_t e, pad;
};

// exp(d) as mantissa/exponent for any finite d (|d| up to ~1e9).
HB_HD RhoME rho_from_logratio(double d)
{
    RhoME r;
    const double L2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01,
                 LN2_LO = 1.90821492927058770002e-10;
    d = d > 7e8 ? 7e8 : (d < -7e8 ? -7e8 : d);
    double nf = rint(d * L2E);
    double rem = dfma(nf, -LN2_LO, dfma(nf, -LN2_HI, d));
    r.m = exp(rem);
    r.e = (int32_t)nf;
    r.pad = 0;
    return r;
}

struct Binom2Const {
    double prob[2], q[2];       // p_k, 1 - p_k
    double logp[2], logq[2];    // host log(p_k), log(1 - p_k)
    double logPi[4], logdelta[2];
    double Pi[4], delta[2];
    int32_t bb; // emissions are beta-binomial: only the table is valid, counts above it are an error
};

struct Binom2Args {
    const int32_t *x, *n; // x[t*ld+b] successes, n[t*ld+b] trials
    int64_t ld, T, B;
    int32_t tiled; // as Norm3Args::tiled: warp-tiled layout of x, n, y, gamma planes and the scratch
    int32_t packed; // 1: x holds both counts per position, (n << 16) | x (each < 65536); n is unused
    int64_t gplane, nchunks;
    const int64_t *seg_off, *chunk_off;
    const int32_t *seg_order;
    int32_t nseg, nblk_per_seg;
    const double2 *lut_lb;  // [(n(n+1)/2 + x)] -> (logb_0, logb_1)
    const RhoME *lut_rho;   //                  -> b_0/b_1 = m * 2^e
    const double *lut_rho_d; //                 -> b_0/b_1 as a double, < 0 = beyond 2^+-900 (hb_binom2_fast.cuh)
    int32_t lut_nmax;
    // optional: (logb_0, logb_1) of every position evaluated beforehand by binom2_emit_kernel,
    // pre_lb[t*ld+b] (pooled chains: deep counts at every position, the saddle-point evaluation is
    // taken off the sequential chain).  Used by the Viterbi-only general kernel.
    const double2 *pre_lb;
    uint8_t *y;
    int64_t ldy;
    double *gamma;
    int64_t ldg;
    double *ll;
    uint64_t *psi; // [nchunks][B]  (2 bits / step used)
    double *ckpt;  // [nchunks][2][B]
    int32_t *status;
    int32_t *gate;          // lean kernel: set to 1 when the launch needs the general kernel
    const int32_t *gate_in; // general kernel: run only if NULL or *gate_in != 0
    Binom2Const c;
};

// ---- nmath restatement for the device slow path (n > lut_nmax) --------------------------------
HB_HD_NOINLINE double dev_stirlerr(double n)
{
    const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778,
                 S2 = 0.00079365079365079365079365, S3 = 0.000595238095238095238095238,
                 S4 = 0.0008417508417508417508417508;
    const double sferr[16] = {0.0,
                              0.0810614667953272582196702,
                              0.0413406959554092940938221,
                              0.02767792568499833914878929,
                              0.02079067210376509311152277,
                              0.01664469118982119216319487,
                              0.01387612882307074799874573,
                              0.01189670994589177009505572,
                              0.010411265261972096497478567,
                              0.009255462182712732917728637,
                              0.008330563433362871256469318,
                              0.007573675487951840794972024,
                              0.006942840107209529865664152,
                              0.006408994188004207068439631,
                              0.005951370112758847735624416,
                              0.005554733551962801371038690};
    if (n <= 15.0) { // integer n only on this path; selects instead of an indexed local array
        const int k = (int)n;
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i++) v = (k == i) ? sferr[i] : v;
        return v;
    }
    double nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

HB_HD_NOINLINE double dev_bd0(double x, double np)
{
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < 2.2250738585072014e-308) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

HB_HD_NOINLINE double dev_dbinom_log(double x, double n, double p, double q, double logp,
                                     double logq)
{
    if (p == 0) return (x == 0) ? 0. : -INFINITY;
    if (q == 0) return (x == n) ? 0. : -INFINITY;
    if (x == 0) {
        if (n == 0) return 0.;
        return (p < 0.1) ? -dev_bd0(n, n * q) - n * p : n * logq;
    }
    if (x == n) return (q < 0.1) ? -dev_bd0(n, n * p) - n * q : n * logp;
    if (x < 0 || x > n) return -INFINITY;
    double lc = dev_stirlerr(n) - dev_stirlerr(x) - dev_stirlerr(n - x) - dev_bd0(x, n * p) -
                dev_bd0(n - x, n * q);
    double lf = 1.837877066409345483560659472811 + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

HB_HD_NOINLINE void binom2_emit_slow(const Binom2Const &c, int32_t xv, int32_t nv, double &lb0,
                                     double &lb1, RhoME &rho)
{
    lb0 = dev_dbinom_log((double)xv, (double)nv, c.prob[0], c.q[0], c.logp[0], c.logq[0]);
    lb1 = dev_dbinom_log((double)xv, (double)nv, c.prob[1], c.q[1], c.logp[1], c.logq[1]);
    if (lb0 == -INFINITY && lb1 == -INFINITY) {
        rho.m = 1.0;
        rho.e = 0;
    } else {
        rho = rho_from_logratio(lb0 - lb1);
    }
}

// Emission of one step: (lb0, lb1) exact-path values and rho = b_0/b_1.
HB_HD void binom2_emit(const Binom2Args &a, int32_t xv, int32_t nv, double &lb0, double &lb1,
                       RhoME &rho, bool want_lb, bool want_rho)
{
    if (nv >= 0 && nv <= a.lut_nmax && xv >= 0 && xv <= nv) {
        int64_t ix = (int64_t)nv * (nv + 1) / 2 + xv;
        if (want_lb) {
            double2 v = a.lut_lb[ix];
            lb0 = v.x;
            lb1 = v.y;
        }
        if (want_rho) rho = a.lut_rho[ix];
    } else {
        if (a.c.bb) *a.status = 2;
        binom2_emit_slow(a.c, xv, nv, lb0, lb1, rho);
    }
}

HB_HD uint32_t binom2_vit_step(const Binom2Const &c, double &n0, double &n1, double l0, double l1)
{
    double c00 = n0 + c.logPi[0], c10 = n1 + c.logPi[2];
    double c01 = n0 + c.logPi[1], c11 = n1 + c.logPi[3];
    uint32_t p0 = 0, p1 = 0;
    double m0 = c00, m1 = c01;
    if (c10 > m0) { m0 = c10; p0 = 1; }
    if (c11 > m1) { m1 = c11; p1 = 1; }
    n0 = m0 + l0;
    n1 = m1 + l1;
    return p0 | (p1 << 1);
}

// ---- forward-backward in the linear domain with a RELATIVE binary exponent ----------------------
// Binomial evidence at one position can be hundreds of nats (deep coverage, pooled groups), far
// beyond what a pair of doubles can hold side by side.  The two filter components are therefore
// kept as mantissas in [1,2) (or 0) plus the integer d = log2 scale of state 0 relative to state 1:
//     phi_0 = m0 * 2^d,   phi_1 = m1          (the common scale goes to the exponent sum `esum`)
// and the emission ratio as rho = mr * 2^er.  Components further apart than 2^1000 stop
// interacting (error < 1e-300); nothing saturates.
#define HB_DCLAMP 1000
HB_HD int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Rel2 {
    double m0, m1;
    int32_t d;
};

// Bring (m0 * 2^d, m1) to a common scale: a0, a1 doubles, returns the common exponent C = max(d, 0).
HB_HD int32_t rel2_align(double m0, double m1, int32_t d, double &a0, double &a1)
{
    int32_t sh = clampi(d, -HB_DCLAMP, HB_DCLAMP);
    a0 = m0 * pow2i(sh < 0 ? sh : 0);
    a1 = m1 * pow2i(sh > 0 ? -sh : 0);
    return d > 0 ? d : 0;
}

// Renormalise (q0 * 2^c0, q1) -> mantissas + relative exponent; returns the exponent taken out of state 1.
HB_HD int32_t rel2_norm(double q0, int32_t c0, double q1, Rel2 &r)
{
    int32_t x0 = q0 > 0 ? expo(q0) : 0;
    int32_t x1 = q1 > 0 ? expo(q1) : 0;
    r.m0 = q0 * pow2i(-x0);
    r.m1 = q1 * pow2i(-x1);
    int64_t dd = (int64_t)x0 + c0 - x1;
    r.d = q0 > 0 ? (q1 > 0 ? (int32_t)(dd > (1 << 30) ? (1 << 30) : (dd < -(1 << 30) ? -(1 << 30) : dd))
                           : (1 << 30))
                 : -(1 << 30);
    return x1;
}

// phi <- (phi Pi) . (rho, 1).  FIRST: phi = delta . (rho, 1).  Returns the exponent moved to esum.
template <bool FIRST>
HB_HD int64_t binom2_fwd_step(const Binom2Const &c, Rel2 &f, double mr, int32_t er)
{
    double p0, p1;
    int32_t C = 0;
    if (FIRST) {
        p0 = c.delta[0];
        p1 = c.delta[1];
    } else {
        double a0, a1;
        C = rel2_align(f.m0, f.m1, f.d, a0, a1);
        p0 = dfma(a1, c.Pi[2], a0 * c.Pi[0]);
        p1 = dfma(a1, c.Pi[3], a0 * c.Pi[1]);
    }
    int32_t x1 = rel2_norm(p0 * mr, er, p1, f);
    return (int64_t)C + x1;
}

// beta <- Pi ((rho, 1) . beta)
HB_HD void binom2_bwd_step(const Binom2Const &c, Rel2 &b, double mr, int32_t er)
{
    double a0, a1;
    int64_t dd = (int64_t)b.d + er;
    int32_t d = (int32_t)(dd > (1 << 30) ? (1 << 30) : (dd < -(1 << 30) ? -(1 << 30) : dd));
    rel2_align(b.m0 * mr, b.m1, d, a0, a1);
    double q0 = dfma(c.Pi[1], a1, c.Pi[0] * a0);
    double q1 = dfma(c.Pi[3], a1, c.Pi[2] * a0);
    rel2_norm(q0, 0, q1, b);
}

// counts of one position: separate planes, or one packed word
HB_HD void load_counts(const Binom2Args &a, const int32_t *xp, const int32_t *np, int64_t o, int32_t &xv,
                       int32_t &nv)
{
    const int32_t w = xp[o];
    if (a.packed) {
        xv = w & 0xffff;
        nv = (int32_t)((uint32_t)w >> 16);
    } else {
        xv = w;
        nv = np[o];
    }
}

template <bool VIT, bool FB>
HB_HD void binom2_chain(const Binom2Args &a, int32_t seg, int64_t b, double *sm, int32_t sms)
{
    const Binom2Const &c = a.c;
    const int64_t t0 = a.seg_off[seg];
    const int64_t L = a.seg_off[seg + 1] - t0;
    if (L <= 0) return;
    const int64_t nch = (L + HB_CHUNK - 1) / HB_CHUNK;
    const int64_t cbase = a.chunk_off[seg];
    const int64_t ldx = lay_stride(a.tiled, a.ld), sB = lay_stride(a.tiled, a.B);
    const int32_t *xp = a.x + lay_off(a.tiled, a.T, a.ld, t0, b);
    const int32_t *np = a.packed ? nullptr : a.n + lay_off(a.tiled, a.T, a.ld, t0, b);
    uint64_t *psi0 = a.psi + lay_off(a.tiled, a.nchunks, a.B, cbase, b);
    double *ckpt0 = a.ckpt + (a.tiled ? (((b >> 5) * a.nchunks + cbase) * 3 << 5) + (b & 31) : cbase * 3 * a.B + b);

    double n0 = 0, n1 = 0, lsum = 0;
    Rel2 f;
    f.m0 = f.m1 = 0;
    f.d = 0;
    int64_t esum = 0;
    const bool want_ll = FB && a.ll != nullptr;

    int32_t xc[HB_CHUNK], nc[HB_CHUNK], xn[HB_CHUNK], nn[HB_CHUNK];
    // pooled chains: pre-evaluated emissions, fetched one chunk ahead (+ L2 prefetch further ahead):
    // only a handful of threads run, so nothing else hides the latency of these loads
    const bool pre = !FB && a.pre_lb != nullptr;
    const double2 *plb = pre ? a.pre_lb + lay_off(a.tiled, a.T, a.ld, t0, b) : nullptr;
    double2 lbc[HB_CHUNK], lbn[HB_CHUNK];
#pragma unroll
    for (int j = 0; j < HB_CHUNK; j++) {
        lbc[j].x = lbc[j].y = lbn[j].x = lbn[j].y = 0.0;
        if (pre) {
            if (j < L) lbc[j] = plb[(int64_t)j * ldx];
            xc[j] = nc[j] = 0;
        } else {
            xc[j] = nc[j] = 0;
            if (j < L) load_counts(a, xp, np, (int64_t)j * ldx, xc[j], nc[j]);
        }
    }
    for (int64_t ch = 0; ch < nch; ch++) {
        const int64_t i0 = ch * HB_CHUNK;
        {
            const int64_t inext = i0 + HB_CHUNK;
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++) {
                if (pre) {
                    if (inext + j < L) lbn[j] = plb[(inext + j) * ldx];
                    if (inext + 4 * HB_CHUNK + j < L) prefetch_l2(plb + (inext + 4 * HB_CHUNK + j) * ldx);
                    xn[j] = nn[j] = 0;
                } else {
                    xn[j] = nn[j] = 0;
                    if (inext + j < L) load_counts(a, xp, np, (inext + j) * ldx, xn[j], nn[j]);
                }
            }
        }
        uint64_t pk = 0;
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) {
            if (i0 + j < L) {
                double l0 = 0, l1 = 0;
                RhoME rho;
                rho.m = 1.0;
                rho.e = 0;
                if (pre) {
                    l0 = lbc[j].x;
                    l1 = lbc[j].y;
                } else {
                    binom2_emit(a, xc[j], nc[j], l0, l1, rho, VIT || want_ll, FB);
                }
                if (VIT) {
                    if (i0 + j == 0) {
                        n0 = c.logdelta[0] + l0;
                        n1 = c.logdelta[1] + l1;
                    } else {
                        uint32_t ps = binom2_vit_step(c, n0, n1, l0, l1);
                        pk |= (uint64_t)ps << (2 * j);
                    }
                }
                if (FB) {
                    if (i0 + j == 0)
                        esum += binom2_fwd_step<true>(c, f, rho.m, rho.e);
                    else
                        esum += binom2_fwd_step<false>(c, f, rho.m, rho.e);
                    lsum += l1;
                }
            }
        }
        if (VIT) psi0[ch * sB] = pk;
        if (FB && ch + 1 < nch) {
            double *ck = ckpt0 + (ch + 1) * 3 * sB;
            ck[0] = f.m0;
            ck[sB] = f.m1;
            ck[2 * sB] = mk64(0, f.d);
        }
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) {
            xc[j] = xn[j];
            nc[j] = nn[j];
            lbc[j] = lbn[j];
        }
    }

    if (want_ll) {
        // LL = log(phi_0 + phi_1) + ln2 * esum + sum_t log b_1
        const double LN2 = 0.693147180559945309417232121458;
        double a0, a1;
        int32_t C = rel2_align(f.m0, f.m1, f.d, a0, a1);
        a.ll[(int64_t)seg * a.B + b] = (log(a0 + a1) + LN2 * (double)(esum + C)) + lsum;
        if (!VIT && !a.gamma) return; // HB_WANT_LL alone: the forward sweep is the whole job
    }

    int32_t ycur = 0;
    if (VIT) {
        if (n0 == -INFINITY || n1 == -INFINITY) *a.status = 1;
        if (n1 > n0) ycur = 1;
    }
    Rel2 bt;
    bt.m0 = bt.m1 = 1.0;
    bt.d = 0;
    if (FB) {
        const int64_t i0 = (nch - 1) * HB_CHUNK;
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) {
            xc[j] = nc[j] = 0;
            if (i0 + j < L) load_counts(a, xp, np, (i0 + j) * ldx, xc[j], nc[j]);
        }
    }
    for (int64_t ch = nch - 1; ch >= 0; ch--) {
        const int64_t i0 = ch * HB_CHUNK;
        if (FB) {
            const int64_t iprev = i0 - HB_CHUNK;
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++) {
                xn[j] = nn[j] = 0;
                if (iprev >= 0) load_counts(a, xp, np, (iprev + j) * ldx, xn[j], nn[j]);
            }
        }
        uint64_t pk = 0;
        if (VIT) {
            pk = psi0[ch * sB];
            if (ch >= 6) prefetch_l2(psi0 + (ch - 6) * sB);
        }
        if (FB) {
            if (ch > 0) {
                const double *ck = ckpt0 + ch * 3 * sB;
                f.m0 = ck[0];
                f.m1 = ck[sB];
                f.d = lo32(ck[2 * sB]);
            }
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++) {
                if (i0 + j < L) {
                    double l0, l1;
                    RhoME rho;
                    rho.m = 1.0;
                    rho.e = 0;
                    binom2_emit(a, xc[j], nc[j], l0, l1, rho, false, true);
                    if (i0 + j == 0)
                        binom2_fwd_step<true>(c, f, rho.m, rho.e);
                    else
                        binom2_fwd_step<false>(c, f, rho.m, rho.e);
                    double *s = sm + (int64_t)j * HB_BINOM2_SLOT * sms;
                    s[0] = f.m0;
                    s[sms] = f.m1;
                    s[2 * sms] = mk64(f.d, rho.e);
                    s[3 * sms] = rho.m;
                }
            }
        }
#pragma unroll
        for (int j = HB_CHUNK - 1; j >= 0; j--) {
            const int64_t i = i0 + j;
            if (i < L) {
                if (VIT) {
                    a.y[lay_off(a.tiled, a.T, a.ldy, t0 + i, b)] = (uint8_t)(ycur + 1);
                    ycur = (int32_t)((pk >> (2 * j + ycur)) & 1u);
                }
                if (FB) {
                    const double *s = sm + (int64_t)j * HB_BINOM2_SLOT * sms;
                    const double pkd = s[2 * sms];
                    const int32_t fd = hi32(pkd), er = lo32(pkd);
                    // gamma_t ~ (m0 b0 2^(d_f + d_b), m1 b1)
                    int64_t dd = (int64_t)fd + bt.d;
                    int32_t d = (int32_t)(dd > (1 << 30) ? (1 << 30) : (dd < -(1 << 30) ? -(1 << 30) : dd));
                    double g0, g1;
                    rel2_align(s[0] * bt.m0, s[sms] * bt.m1, d, g0, g1);
                    double rs = rcp_fast(g0 + g1);
                    double *gp = a.gamma + lay_off(a.tiled, a.T, a.ldg, t0 + i, b);
                    gp[0] = g0 * rs;
                    gp[a.gplane] = g1 * rs;
                    binom2_bwd_step(c, bt, s[3 * sms], er);
                }
            }
        }
        if (FB) {
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++) {
                xc[j] = xn[j];
                nc[j] = nn[j];
            }
        }
    }
}

} // namespace hb
