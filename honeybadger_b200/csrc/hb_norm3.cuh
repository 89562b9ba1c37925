// hb_norm3.cuh — 3-state Gaussian chain: fused Viterbi + forward-backward, one thread per chain.
//
// Replaces HiddenMarkov::dthmm(..., "norm", ...) -> HiddenMarkov::Viterbi at
// R/HoneyBADGER.R:1150-1151 and R/HoneyBADGER_gexp.R:503-504, and adds the posteriors of
// HiddenMarkov::forwardback.
//
// Data flow per chain (cell column b, segment s), gene-major x[t*ld + b]:
//   forward sweep   t = 0..L-1 : exact log-space Viterbi recursion (back-pointers packed 6 bit /
//                                step, 8 steps per uint64 -> psi scratch) and the scaled linear
//                                forward filter (checkpoint of phi every 8 steps -> ckpt scratch)
//   backward sweep  chunks of 8, last to first: replay the forward filter inside the chunk from its
//                                checkpoint into per-thread shared-memory slots, then run beta
//                                backwards through the chunk: gamma_t = phi_t . beta_t (normalised),
//                                and follow the back-pointers: y_t.
// HBM traffic per step: read x twice (8+8), psi 1+1, checkpoints 3+3, write gamma 24 + y 1.
//
// Numerics
//   Viterbi is bit-exact w.r.t. the reference operation order: logb_k = -((C + (0.5 z_k) z_k) + log sd)
//   with z_k = RN((x - mean_k)/sd) (correctly rounded via div_by_const), one rounded add per
//   nu_j + logPi[j][k], exact max with the LOWEST index winning ties, one rounded add of logb.
//   log(sd), log(Pi), log(delta) come from the caller (host libm) so they match R's.  Build with
//   -fmad=false: no contraction anywhere on this path.
//   Forward-backward runs in the linear domain on emission RATIOS b_k/b_1 = exp(a_k x + c_k) (the
//   posteriors are invariant to the common factor b_1), rescaled by an exact power of two every
//   step; the log-likelihood adds the removed exponents and sum log b_1 back.
#pragma once
#include "hb_device.cuh"

namespace hb {

#define HB_CHUNK 8
#define HB_NORM3_SLOT 5 // doubles per step kept for the backward sweep: phi[3], E0, E2

struct Norm3Const {
    double mean[3];
    double logPi[9]; // row-major j->k, host-computed log(Pi)
    double logdelta[3];
    double Pi[9];
    double delta[3];
    double la, lt;  // SYM: log(diag), log(off-diag)
    double amt, tt; // SYM: (diag - offdiag), offdiag
};

struct Norm3Args {
    const double *x;
    int64_t ld, T, B;
    const int64_t *seg_off;   // device [nseg+1]
    const int64_t *chunk_off; // device [nseg+1]: prefix sum of ceil(L_s / HB_CHUNK)
    const int32_t *seg_order; // device [nseg]: segments by decreasing length
    int32_t nseg, nblk_per_seg;
    const double *sd, *inv_sd, *log_sd; // device, [B] or [nseg*B]
    int32_t sd_per_seg;
    uint8_t *y;
    int64_t ldy;
    double *gamma;
    int64_t ldg; // gamma[(k*T + t)*ldg + b]
    double *ll;  // [nseg*B]
    uint64_t *psi; // [nchunks][B]
    double *ckpt;  // [nchunks][3][B]
    int32_t *status; // set to 1 when a chain ends with nu == -inf (R: "Problems With Underflow")
    Norm3Const c;
};

#define HB_LN_SQRT_2PI 0.918938533204672741780329736406

// log dnorm(x, mean_k, sd) for the three states, R nmath operation order.
HB_HD void norm3_logb(double xv, const Norm3Const &c, double sdv, double isd, double lsd,
                      double &l0, double &l1, double &l2, double &h1)
{
    double z0 = div_by_const(xv - c.mean[0], sdv, isd);
    double z1 = div_by_const(xv - c.mean[1], sdv, isd);
    double z2 = div_by_const(xv - c.mean[2], sdv, isd);
    double h0 = (0.5 * z0) * z0;
    h1 = (0.5 * z1) * z1;
    double h2 = (0.5 * z2) * z2;
    l0 = -((HB_LN_SQRT_2PI + h0) + lsd);
    l1 = -((HB_LN_SQRT_2PI + h1) + lsd);
    l2 = -((HB_LN_SQRT_2PI + h2) + lsd);
}

// One Viterbi step; returns the three back-pointers packed 2 bits each (k=0 in bits 0-1).
template <bool SYM>
HB_HD uint32_t norm3_vit_step(const Norm3Const &c, double &n0, double &n1, double &n2, double l0,
                              double l1, double l2)
{
    double c00, c10, c20, c01, c11, c21, c02, c12, c22; // c_jk = nu_j + logPi[j][k]
    if (SYM) {
        double A0 = n0 + c.la, A1 = n1 + c.la, A2 = n2 + c.la;
        double B0 = n0 + c.lt, B1 = n1 + c.lt, B2 = n2 + c.lt;
        c00 = A0; c10 = B1; c20 = B2;
        c01 = B0; c11 = A1; c21 = B2;
        c02 = B0; c12 = B1; c22 = A2;
    } else {
        c00 = n0 + c.logPi[0]; c10 = n1 + c.logPi[3]; c20 = n2 + c.logPi[6];
        c01 = n0 + c.logPi[1]; c11 = n1 + c.logPi[4]; c21 = n2 + c.logPi[7];
        c02 = n0 + c.logPi[2]; c12 = n1 + c.logPi[5]; c22 = n2 + c.logPi[8];
    }
    uint32_t p0 = 0, p1 = 0, p2 = 0;
    double m0 = c00, m1 = c01, m2 = c02;
    if (c10 > m0) { m0 = c10; p0 = 1; }
    if (c20 > m0) { m0 = c20; p0 = 2; }
    if (c11 > m1) { m1 = c11; p1 = 1; }
    if (c21 > m1) { m1 = c21; p1 = 2; }
    if (c12 > m2) { m2 = c12; p2 = 1; }
    if (c22 > m2) { m2 = c22; p2 = 2; }
    n0 = m0 + l0;
    n1 = m1 + l1;
    n2 = m2 + l2;
    return p0 | (p1 << 2) | (p2 << 4);
}

// Per-chain constants of the linear-domain filter.
struct Norm3Fb {
    // SYM: rho_0 = K*exp(-u), rho_2 = K*exp(u), u = A*x + Bc; aK = K*(a-t), tK = K*t
    double A, Bc, K, aK, tK;
    // general: rho_k = exp(a_k x + c_k), k = 0, 2
    double a0, c0, a2, c2;
};

template <bool SYM>
HB_HD Norm3Fb norm3_fb_consts(const Norm3Const &c, double isd)
{
    Norm3Fb k;
    if (SYM) {
        double m = c.mean[2] - c.mean[1];
        double mi = m * isd;
        k.A = mi * isd;
        k.Bc = -k.A * c.mean[1];
        k.K = exp(-0.5 * mi * mi);
        k.aK = k.K * c.amt;
        k.tK = k.K * c.tt;
        k.a0 = k.c0 = k.a2 = k.c2 = 0;
    } else {
        double i2 = isd * isd;
        k.a0 = (c.mean[0] - c.mean[1]) * i2;
        k.c0 = -0.5 * (c.mean[0] * c.mean[0] - c.mean[1] * c.mean[1]) * i2;
        k.a2 = (c.mean[2] - c.mean[1]) * i2;
        k.c2 = -0.5 * (c.mean[2] * c.mean[2] - c.mean[1] * c.mean[1]) * i2;
        k.A = k.Bc = 0;
        k.K = 1.0;
        k.aK = k.tK = 0;
    }
    return k;
}

// Emission ratios of one step, without the 2^n factors applied: E0 = r0 * 2^n0, E2 = r2 * 2^n2.
template <bool SYM>
HB_HD void norm3_ratios(const Norm3Fb &k, double xv, double &r0, int32_t &n0, double &r2,
                        int32_t &n2)
{
    if (SYM) {
        double u = dfma(xv, k.A, k.Bc);
        int32_t n;
        exp_pair(u, r2, r0, n);
        n2 = n;
        n0 = -n;
    } else {
        exp_one(dfma(xv, k.a0, k.c0), r0, n0);
        exp_one(dfma(xv, k.a2, k.c2), r2, n2);
    }
}

// Forward filter step: phi <- ((phi Pi) . rho) * 2^-es with es = exponent of sum(phi_in).
// FIRST: phi_in is delta and there is no transition.  Returns es (added to the chain's exponent sum).
template <bool SYM, bool FIRST>
HB_HD int32_t norm3_fwd_step(const Norm3Const &c, const Norm3Fb &k, double &f0, double &f1,
                             double &f2, double r0, int32_t n0, double r2, int32_t n2)
{
    double p0, p1, p2;
    int32_t es;
    if (FIRST) {
        p0 = c.delta[0] * k.K;
        p1 = c.delta[1];
        p2 = c.delta[2] * k.K;
        es = 0;
    } else {
        double s = (f0 + f1) + f2;
        es = expo(s);
        if (SYM) {
            double ts = c.tt * s, tKs = k.tK * s;
            p1 = dfma(c.amt, f1, ts);
            p0 = dfma(k.aK, f0, tKs);
            p2 = dfma(k.aK, f2, tKs);
        } else {
            p0 = dfma(f2, c.Pi[6], dfma(f1, c.Pi[3], f0 * c.Pi[0]));
            p1 = dfma(f2, c.Pi[7], dfma(f1, c.Pi[4], f0 * c.Pi[1]));
            p2 = dfma(f2, c.Pi[8], dfma(f1, c.Pi[5], f0 * c.Pi[2]));
        }
    }
    f1 = p1 * pow2i(-es);
    f0 = p0 * scale2(r0, n0 - es);
    f2 = p2 * scale2(r2, n2 - es);
    return es;
}

// Backward step: beta <- Pi (rho . beta) * 2^-eb (eb = exponent of the previous step's sum);
// updates eb with this step's sum.
template <bool SYM>
HB_HD void norm3_bwd_step(const Norm3Const &c, const Norm3Fb &k, double &b0, double &b1,
                          double &b2, int32_t &eb, double E0, double E2)
{
    double v0 = scale2(E0, -eb) * b0;
    double v2 = scale2(E2, -eb) * b2;
    double v1 = b1 * pow2i(-eb);
    if (SYM) {
        double sw = dfma(k.K, v0 + v2, v1);
        double tsw = c.tt * sw;
        b1 = dfma(c.amt, v1, tsw);
        b0 = dfma(k.aK, v0, tsw);
        b2 = dfma(k.aK, v2, tsw);
        eb = expo(sw);
    } else {
        b0 = dfma(c.Pi[2], v2, dfma(c.Pi[1], v1, c.Pi[0] * v0));
        b1 = dfma(c.Pi[5], v2, dfma(c.Pi[4], v1, c.Pi[3] * v0));
        b2 = dfma(c.Pi[8], v2, dfma(c.Pi[7], v1, c.Pi[6] * v0));
        eb = expo((b0 + b1) + b2);
    }
}

// The whole chain.  `sm` points at this thread's first shared-memory word; consecutive slots of the
// same thread are `sms` doubles apart (blockDim.x on the device: conflict-free columns).
template <bool SYM, bool VIT, bool FB>
HB_HD void norm3_chain(const Norm3Args &a, int32_t seg, int64_t b, double *sm, int32_t sms)
{
    const Norm3Const &c = a.c;
    const int64_t t0 = a.seg_off[seg];
    const int64_t L = a.seg_off[seg + 1] - t0;
    if (L <= 0) return;
    const int64_t nch = (L + HB_CHUNK - 1) / HB_CHUNK;
    const int64_t cbase = a.chunk_off[seg];
    const int64_t pix = a.sd_per_seg ? (int64_t)seg * a.B + b : b;
    const double sdv = a.sd[pix], isd = a.inv_sd[pix], lsd = a.log_sd[pix];
    const double *xp = a.x + t0 * a.ld + b;
    const Norm3Fb k = norm3_fb_consts<SYM>(c, isd);

    double n0 = 0, n1 = 0, n2 = 0; // Viterbi nu
    double f0 = 0, f1 = 0, f2 = 0; // forward filter
    double hsum = 0;               // sum of (0.5 z_1) z_1 for the log-likelihood
    int64_t esum = 0;              // sum of removed binary exponents

    // ------------------------------------------------------------ forward sweep
    double xc[HB_CHUNK], xn[HB_CHUNK];
#pragma unroll
    for (int j = 0; j < HB_CHUNK; j++) xc[j] = (j < L) ? xp[(int64_t)j * a.ld] : 0.0;
    for (int64_t ch = 0; ch < nch; ch++) {
        const int64_t i0 = ch * HB_CHUNK;
        {
            const int64_t inext = i0 + HB_CHUNK;
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++)
                xn[j] = (inext + j < L) ? xp[(inext + j) * a.ld] : 0.0;
        }
        uint64_t pk = 0;
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) {
            if (i0 + j < L) {
                const double xv = xc[j];
                double h1 = 0;
                if (VIT || (FB && a.ll)) {
                    double l0, l1, l2;
                    norm3_logb(xv, c, sdv, isd, lsd, l0, l1, l2, h1);
                    if (VIT) {
                        if (i0 + j == 0) {
                            n0 = c.logdelta[0] + l0;
                            n1 = c.logdelta[1] + l1;
                            n2 = c.logdelta[2] + l2;
                        } else {
                            uint32_t ps = norm3_vit_step<SYM>(c, n0, n1, n2, l0, l1, l2);
                            pk |= (uint64_t)ps << (6 * j);
                        }
                    }
                }
                if (FB) {
                    double r0, r2;
                    int32_t e0, e2;
                    norm3_ratios<SYM>(k, xv, r0, e0, r2, e2);
                    if (i0 + j == 0)
                        esum += norm3_fwd_step<SYM, true>(c, k, f0, f1, f2, r0, e0, r2, e2);
                    else
                        esum += norm3_fwd_step<SYM, false>(c, k, f0, f1, f2, r0, e0, r2, e2);
                    hsum += h1;
                }
            }
        }
        if (VIT) a.psi[(cbase + ch) * a.B + b] = pk;
        if (FB && ch + 1 < nch) {
            double *ck = a.ckpt + (cbase + ch + 1) * 3 * a.B + b;
            ck[0] = f0;
            ck[a.B] = f1;
            ck[2 * a.B] = f2;
        }
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) xc[j] = xn[j];
    }

    if (FB && a.ll) {
        // LL = log sum(alpha_L) = log(sum phi) + ln2 * esum + sum_t log b_1(x_t)
        const double LN2 = 0.693147180559945309417232121458;
        a.ll[(int64_t)seg * a.B + b] =
            ((log((f0 + f1) + f2) + LN2 * (double)esum) - hsum) - (double)L * (HB_LN_SQRT_2PI + lsd);
    }

    // ------------------------------------------------------------ backward sweep
    int32_t ycur = 0;
    if (VIT) {
        if (n0 == -INFINITY || n1 == -INFINITY || n2 == -INFINITY) *a.status = 1;
        double mv = n0;
        if (n1 > mv) { mv = n1; ycur = 1; }
        if (n2 > mv) { mv = n2; ycur = 2; }
    }
    double b0 = 1.0, b1 = 1.0, b2 = 1.0;
    int32_t eb = 0;
    {
        const int64_t i0 = (nch - 1) * HB_CHUNK;
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) xc[j] = (i0 + j < L) ? xp[(i0 + j) * a.ld] : 0.0;
    }
    for (int64_t ch = nch - 1; ch >= 0; ch--) {
        const int64_t i0 = ch * HB_CHUNK;
        if (FB) {
            const int64_t iprev = i0 - HB_CHUNK;
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++)
                xn[j] = (iprev >= 0) ? xp[(iprev + j) * a.ld] : 0.0;
        }
        uint64_t pk = 0;
        if (VIT) pk = a.psi[(cbase + ch) * a.B + b];
        if (FB) {
            // replay the filter through the chunk from its checkpoint, keeping phi and the ratios
            if (ch > 0) {
                const double *ck = a.ckpt + (cbase + ch) * 3 * a.B + b;
                f0 = ck[0];
                f1 = ck[a.B];
                f2 = ck[2 * a.B];
            }
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++) {
                if (i0 + j < L) {
                    double r0, r2;
                    int32_t e0, e2;
                    norm3_ratios<SYM>(k, xc[j], r0, e0, r2, e2);
                    if (i0 + j == 0)
                        norm3_fwd_step<SYM, true>(c, k, f0, f1, f2, r0, e0, r2, e2);
                    else
                        norm3_fwd_step<SYM, false>(c, k, f0, f1, f2, r0, e0, r2, e2);
                    double *s = sm + (int64_t)j * HB_NORM3_SLOT * sms;
                    s[0] = f0;
                    s[sms] = f1;
                    s[2 * sms] = f2;
                    s[3 * sms] = scale2(r0, e0);
                    s[4 * sms] = scale2(r2, e2);
                }
            }
        }
#pragma unroll
        for (int j = HB_CHUNK - 1; j >= 0; j--) {
            const int64_t i = i0 + j;
            if (i < L) {
                if (VIT) {
                    a.y[(t0 + i) * a.ldy + b] = (uint8_t)(ycur + 1);
                    ycur = (int32_t)((pk >> (6 * j + 2 * ycur)) & 3u);
                }
                if (FB) {
                    const double *s = sm + (int64_t)j * HB_NORM3_SLOT * sms;
                    double g0 = s[0] * b0, g1 = s[sms] * b1, g2 = s[2 * sms] * b2;
                    double rs = rcp_fast((g0 + g1) + g2);
                    double *gp = a.gamma + (t0 + i) * a.ldg + b;
                    const int64_t plane = a.T * a.ldg;
                    gp[0] = g0 * rs;
                    gp[plane] = g1 * rs;
                    gp[2 * plane] = g2 * rs;
                    norm3_bwd_step<SYM>(c, k, b0, b1, b2, eb, s[3 * sms], s[4 * sms]);
                }
            }
        }
        if (FB) {
#pragma unroll
            for (int j = 0; j < HB_CHUNK; j++) xc[j] = xn[j];
        }
    }
}

} // namespace hb
