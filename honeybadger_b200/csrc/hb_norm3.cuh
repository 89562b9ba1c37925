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
//
// Instruction economy (the kernel is fp64-issue bound, not HBM bound — see DESIGN.md): every
// floating constant that is not a short immediate lives in the kernel-parameter block
// (Norm3Const::k) so it is fetched from the constant bank instead of being rebuilt from two 32-bit
// immediates per use; full 8-step chunks run without bounds checks; the next chunk's x is
// prefetched into the register the current step just consumed.
#pragma once
#include "hb_device.cuh"

namespace hb {

// steps per chunk: one back-pointer word and one checkpoint per chunk.  (10 was tried: 3 % less
// scratch traffic, no gain in time on either chain family, so the kernels assume 8.)
#define HB_CHUNK 8

// indices into Norm3Const::k
enum {
    HBK_LNSQRT2PI = 0,
    HBK_L2E,
    HBK_LN2HI,
    HBK_LN2LO,
    HBK_E12,
    HBK_E10,
    HBK_E8,
    HBK_E6,
    HBK_E4,
    HBK_O13,
    HBK_O11,
    HBK_O9,
    HBK_O7,
    HBK_O5,
    HBK_O3,
    HBK_LN2,
    HBK_COUNT
};

struct Norm3Const {
    double mean[3];
    double logPi[9]; // row-major j->k, host-computed log(Pi)
    double logdelta[3];
    double Pi[9];
    double delta[3];
    double la, lt;  // SYM: log(diag), log(off-diag)
    double amt, tt; // SYM: (diag - offdiag), offdiag
    // speculate-and-verify path (hb_norm3_fast.cuh)
    double tau;       // lt - la
    double dl0, dl2;  // logdelta[k] - logdelta[1]
    double mumax2;    // max_k mean[k]^2
    double dmu2;      // max_k (mean[k] - mean[1])^2
    double ldabs;     // max finite |logdelta[k]|
    double eps_scale; // 1 (or +Inf: test hook that sends every chain through the exact fallback)
    double k[HBK_COUNT];
};

// Fills the numeric constants (host side; the same values on every platform).
inline void norm3_fill_k(Norm3Const &c)
{
    c.k[HBK_LNSQRT2PI] = 0.918938533204672741780329736406;
    c.k[HBK_L2E] = 1.4426950408889634074;
    c.k[HBK_LN2HI] = 0.69314670562744140625;     // 0x3FE62E4200000000: 21 significant bits
    c.k[HBK_LN2LO] = 4.74932503903167232121458e-07; // ln2 - LN2HI, rounded once
    c.k[HBK_E12] = 1.0 / 479001600.0;
    c.k[HBK_E10] = 1.0 / 3628800.0;
    c.k[HBK_E8] = 1.0 / 40320.0;
    c.k[HBK_E6] = 1.0 / 720.0;
    c.k[HBK_E4] = 1.0 / 24.0;
    c.k[HBK_O13] = 1.0 / 6227020800.0;
    c.k[HBK_O11] = 1.0 / 39916800.0;
    c.k[HBK_O9] = 1.0 / 362880.0;
    c.k[HBK_O7] = 1.0 / 5040.0;
    c.k[HBK_O5] = 1.0 / 120.0;
    c.k[HBK_O3] = 1.0 / 6.0;
    c.k[HBK_LN2] = 0.693147180559945309417232121458;
}

struct Norm3Args {
    const double *x;
    int64_t ld, T, B;
    // layout: 0 = gene-major x[t*ld + b]; 1 = warp-tiled x[((b/32)*T + t)*32 + b%32] (every warp's
    // rows are contiguous: sequential 256-byte accesses per warp, 2 KB per chunk), also for y, each
    // gamma plane (plane stride gplane) and the scratch (psi, ckpt)
    int32_t tiled;
    int64_t gplane, nchunks;
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
    int32_t *status; // [0] set to 1 when a chain ends with nu == -inf (R: "Problems With Underflow")
                     // [1] chains of the speculative path that were re-run exactly
    int64_t *flagged;     // [flagged_cap] seg * B + b of the first re-run chains (diagnostics / parity tests)
    int32_t flagged_cap;
    Norm3Const c;
};

// exp(u) and exp(-u): Pp = e^r, Pm = e^-r, u = n ln2 + r (see hb_device.cuh exp_pair; this variant
// reads its constants from the parameter block).
HB_HD void norm3_exp_pair(const Norm3Const &c, double u, double &Pp, double &Pm, int32_t &n)
{
    const double MAGIC = 6755399441055744.0; // 1.5 * 2^52 (short immediate)
    double t = dfma(u, c.k[HBK_L2E], MAGIC);
    int32_t ni = lo32(t);
    double nf = t - MAGIC;
    double r = dfma(nf, -c.k[HBK_LN2HI], u);
    r = dfma(nf, -c.k[HBK_LN2LO], r);
    double q = r * r;
    double ev = dfma(q, c.k[HBK_E12], c.k[HBK_E10]);
    double od = dfma(q, c.k[HBK_O13], c.k[HBK_O11]);
    ev = dfma(q, ev, c.k[HBK_E8]);
    od = dfma(q, od, c.k[HBK_O9]);
    ev = dfma(q, ev, c.k[HBK_E6]);
    od = dfma(q, od, c.k[HBK_O7]);
    ev = dfma(q, ev, c.k[HBK_E4]);
    od = dfma(q, od, c.k[HBK_O5]);
    ev = dfma(q, ev, 0.5);
    od = dfma(q, od, c.k[HBK_O3]);
    ev = dfma(q, ev, 1.0);
    od = dfma(q, od, 1.0);
    od = od * r;
    Pp = ev + od;
    Pm = ev - od;
    n = ni > HB_NCLAMP ? HB_NCLAMP : (ni < -HB_NCLAMP ? -HB_NCLAMP : ni);
}

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
    l0 = -((c.k[HBK_LNSQRT2PI] + h0) + lsd);
    l1 = -((c.k[HBK_LNSQRT2PI] + h1) + lsd);
    l2 = -((c.k[HBK_LNSQRT2PI] + h2) + lsd);
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
    // first index wins ties: strict ">" when a later candidate replaces an earlier one
    const bool a0 = c10 > c00, a1 = c11 > c01, a2 = c12 > c02;
    double m0 = a0 ? c10 : c00, m1 = a1 ? c11 : c01, m2 = a2 ? c12 : c02;
    const bool b0 = c20 > m0, b1 = c21 > m1, b2 = c22 > m2;
    m0 = b0 ? c20 : m0;
    m1 = b1 ? c21 : m1;
    m2 = b2 ? c22 : m2;
    n0 = m0 + l0;
    n1 = m1 + l1;
    n2 = m2 + l2;
    const uint32_t p0 = b0 ? 2u : (a0 ? 1u : 0u);
    const uint32_t p1 = b1 ? 2u : (a1 ? 1u : 0u);
    const uint32_t p2 = b2 ? 2u : (a2 ? 1u : 0u);
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
HB_HD void norm3_ratios(const Norm3Const &c, const Norm3Fb &k, double xv, double &r0, int32_t &n0,
                        double &r2, int32_t &n2)
{
    if (SYM) {
        double u = dfma(xv, k.A, k.Bc);
        int32_t n;
        norm3_exp_pair(c, u, r2, r0, n);
        n2 = n;
        n0 = -n;
    } else {
        double dummy;
        norm3_exp_pair(c, dfma(xv, k.a0, k.c0), r0, dummy, n0);
        norm3_exp_pair(c, dfma(xv, k.a2, k.c2), r2, dummy, n2);
    }
}

// Forward filter step: phi <- ((phi Pi) . rho) * 2^-es with es = exponent of sum(phi_in).
// `first`: phi_in is delta and there is no transition.  Returns es (added to the exponent sum).
template <bool SYM>
HB_HD int32_t norm3_fwd_step(const Norm3Const &c, const Norm3Fb &k, bool first, double &f0,
                             double &f1, double &f2, double r0, int32_t n0, double r2, int32_t n2)
{
    double p0, p1, p2;
    int32_t es;
    if (first) {
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

// Per-chain state carried through both sweeps (kept in registers).
template <bool SYM>
struct Norm3State {
    double sdv, isd, lsd;
    Norm3Fb k;
    double n0, n1, n2; // Viterbi nu
    double f0, f1, f2; // forward filter
    double hsum;       // sum of (0.5 z_1) z_1 (log-likelihood)
    int32_t esum;      // sum of removed binary exponents (|es| <= 502 per step: fine for L < 4e6)
    double b0, b1, b2; // backward
    int32_t eb, ycur;
};

// One forward step (Viterbi recursion + filter).  `first` is the chain's position 0.
template <bool SYM, bool VIT, bool FB, bool LL>
HB_HD uint32_t norm3_forward_one(const Norm3Const &c, Norm3State<SYM> &s, double xv, bool first)
{
    uint32_t ps = 0;
    double h1 = 0;
    if (VIT || (FB && LL)) {
        double l0, l1, l2;
        norm3_logb(xv, c, s.sdv, s.isd, s.lsd, l0, l1, l2, h1);
        if (VIT) {
            if (first) {
                s.n0 = c.logdelta[0] + l0;
                s.n1 = c.logdelta[1] + l1;
                s.n2 = c.logdelta[2] + l2;
            } else {
                ps = norm3_vit_step<SYM>(c, s.n0, s.n1, s.n2, l0, l1, l2);
            }
        }
    }
    if (FB) {
        double r0, r2;
        int32_t e0, e2;
        norm3_ratios<SYM>(c, s.k, xv, r0, e0, r2, e2);
        s.esum += norm3_fwd_step<SYM>(c, s.k, first, s.f0, s.f1, s.f2, r0, e0, r2, e2);
        if (LL) s.hsum += h1;
    }
    return ps;
}

// ---------------------------------------------------------------------------------------------
// Per-thread view of the block's shared memory.  All arrays are [step j][thread]; `st` is the
// element stride between consecutive j for the same thread (blockDim.x on the device, 1 in the
// host emulation), so a warp always touches consecutive addresses (no bank conflicts).
//   va[j] = (phi0, phi1), vb[j] = (phi2, E0), e2[j] = E2 : parked by the replay for the backward pass
//   xs[j]                                               : x staging ring filled by cp.async
struct Norm3Smem {
    double2 *va, *vb;
    double *e2, *xs;
    int32_t *ph; // hb_norm3_fast.cuh: high word of 2^-es per step
    uint8_t *ys; // hb_norm3_fast.cuh, tiled layout: this chain's column of the warp's [HB_CHUNK][32] state tile
    int32_t st;
#if !defined(__CUDA_ARCH__)
    // host emulation of cp.async groups: copies become visible only at wait (catches hazards)
    double pend_v[64];
    double *pend_p[64];
    int pend_g[64];
    int npend = 0, cur_group = 0;
#endif
};
#define HB_NORM3_SMEM_BYTES_PER_THREAD (HB_CHUNK * (16 + 16 + 8 + 8))

HB_HD void xs_refill(Norm3Smem &m, int j, const double *g)
{
#if defined(__CUDA_ARCH__)
    unsigned sa = (unsigned)__cvta_generic_to_shared(m.xs + j * m.st);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g) : "memory");
#else
    m.pend_v[m.npend] = *g;
    m.pend_p[m.npend] = m.xs + j * m.st;
    m.pend_g[m.npend] = m.cur_group;
    m.npend++;
#endif
}
HB_HD void xs_commit(Norm3Smem &m)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    m.cur_group++;
#endif
}
// Wait until at most N of the most recently committed groups are still in flight.
template <int N>
HB_HD void xs_wait(Norm3Smem &m)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#else
    int keep = 0;
    for (int i = 0; i < m.npend; i++) {
        if (m.pend_g[i] < m.cur_group - N) {
            *m.pend_p[i] = m.pend_v[i];
        } else {
            m.pend_v[keep] = m.pend_v[i];
            m.pend_p[keep] = m.pend_p[i];
            m.pend_g[keep] = m.pend_g[i];
            keep++;
        }
    }
    m.npend = keep;
#endif
}

#ifndef HB_UNROLL_N
#define HB_UNROLL_N 4
#endif
#define HB_STR_(x) #x
#define HB_UNROLL_PRAGMA_(n) _Pragma(HB_STR_(unroll n))
#define HB_UNROLL_STEPS HB_UNROLL_PRAGMA_(HB_UNROLL_N)
#define HB_HALF (HB_CHUNK / 2)

// Replay one filter step and park (phi, E0, E2) in slot j for the backward pass.
template <bool SYM>
HB_HD void norm3_replay_one(const Norm3Const &c, Norm3State<SYM> &s, double xv, bool first,
                            Norm3Smem &m, int j)
{
    double r0, r2;
    int32_t e0, e2;
    norm3_ratios<SYM>(c, s.k, xv, r0, e0, r2, e2);
    norm3_fwd_step<SYM>(c, s.k, first, s.f0, s.f1, s.f2, r0, e0, r2, e2);
    double2 a, b;
    a.x = s.f0;
    a.y = s.f1;
    b.x = s.f2;
    b.y = scale2(r0, e0);
    m.va[j * m.st] = a;
    m.vb[j * m.st] = b;
    m.e2[j * m.st] = scale2(r2, e2);
}

// One backward step at global row `row`: Viterbi state, posterior, beta update.
template <bool SYM, bool VIT, bool FB>
HB_HD void norm3_backward_one(const Norm3Args &a, Norm3State<SYM> &s, int64_t row, int64_t b,
                              uint64_t pk, int j, const Norm3Smem &m)
{
    if (VIT) {
        a.y[lay_off(a.tiled, a.T, a.ldy, row, b)] = (uint8_t)(s.ycur + 1);
        s.ycur = (int32_t)((uint32_t)(pk >> (6 * j + 2 * s.ycur)) & 3u);
    }
    if (FB) {
        const double2 va = m.va[j * m.st], vb = m.vb[j * m.st];
        const double E2 = m.e2[j * m.st];
        double g0 = va.x * s.b0, g1 = va.y * s.b1, g2 = vb.x * s.b2;
        double rs = rcp_fast((g0 + g1) + g2);
        double *gp = a.gamma + lay_off(a.tiled, a.T, a.ldg, row, b);
        const int64_t plane = a.gplane;
        gp[0] = g0 * rs;
        gp[plane] = g1 * rs;
        gp[2 * plane] = g2 * rs;
        norm3_bwd_step<SYM>(a.c, s.k, s.b0, s.b1, s.b2, s.eb, vb.y, E2);
    }
}

// The whole chain.
template <bool SYM, bool VIT, bool FB, bool LL>
HB_HD void norm3_chain(const Norm3Args &a, int32_t seg, int64_t b, Norm3Smem &m)
{
    const Norm3Const &c = a.c;
    const int64_t t0 = a.seg_off[seg];
    const int32_t L = (int32_t)(a.seg_off[seg + 1] - t0);
    if (L <= 0) return;
    const int32_t nfull = L / HB_CHUNK;        // chunks with all 8 steps
    const int32_t tail = L - nfull * HB_CHUNK; // steps in the last, partial chunk
    const int64_t cbase = a.chunk_off[seg];
    const int64_t pix = a.sd_per_seg ? (int64_t)seg * a.B + b : b;
    const int64_t ld = lay_stride(a.tiled, a.ld), sB = lay_stride(a.tiled, a.B);
    const double *xp = a.x + lay_off(a.tiled, a.T, a.ld, t0, b);
    uint64_t *psi = a.psi + lay_off(a.tiled, a.nchunks, a.B, cbase, b);
    double *ckpt = a.ckpt + (a.tiled ? (((b >> 5) * a.nchunks + cbase) * 3 << 5) + (b & 31) : cbase * 3 * a.B + b);

    Norm3State<SYM> s;
    s.sdv = a.sd[pix];
    s.isd = a.inv_sd[pix];
    s.lsd = a.log_sd[pix];
    s.k = norm3_fb_consts<SYM>(c, s.isd);
    s.n0 = s.n1 = s.n2 = 0;
    s.f0 = s.f1 = s.f2 = 0;
    s.hsum = 0;
    s.esum = 0;

    // ------------------------------------------------------------ forward sweep
    // x is staged through shared memory by cp.async in half-chunk groups: while a half-chunk is
    // consumed, its slots are refilled with the same positions of the NEXT chunk (distance 8 steps).
    if (nfull > 0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) xs_refill(m, j, xp + (int64_t)j * ld);
            xs_commit(m);
        }
    }
    for (int32_t ch = 0; ch < nfull; ch++) {
        const double *xnext = xp + (int64_t)(ch + 1) * HB_CHUNK * ld;
        const bool more = ch + 1 < nfull;
        uint64_t pk = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            xs_wait<1>(m);
            HB_UNROLL_STEPS
            for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                const double xv = m.xs[j * m.st];
                if (more) xs_refill(m, j, xnext + (int64_t)j * ld);
                uint32_t ps = norm3_forward_one<SYM, VIT, FB, LL>(c, s, xv, j == 0 && ch == 0);
                pk |= (uint64_t)ps << (6 * j);
            }
            xs_commit(m);
        }
        if (VIT) psi[(int64_t)ch * sB] = pk;
        if (FB && (ch + 1 < nfull || tail > 0)) {
            double *ck = ckpt + (int64_t)(ch + 1) * 3 * sB;
            ck[0] = s.f0;
            ck[sB] = s.f1;
            ck[2 * sB] = s.f2;
        }
    }
    xs_wait<0>(m);
    if (tail > 0) { // rolled: at most 7 steps per chain
        uint64_t pk = 0;
        for (int j = 0; j < tail; j++) {
            const double xv = xp[(int64_t)(nfull * HB_CHUNK + j) * ld];
            uint32_t ps = norm3_forward_one<SYM, VIT, FB, LL>(c, s, xv, nfull == 0 && j == 0);
            pk |= (uint64_t)ps << (6 * j);
        }
        if (VIT) psi[(int64_t)nfull * sB] = pk;
    }

    if (FB && LL) {
        // LL = log sum(alpha_L) = log(sum phi) + ln2 * esum + sum_t log b_1(x_t)
        a.ll[(int64_t)seg * a.B + b] = ((log((s.f0 + s.f1) + s.f2) + c.k[HBK_LN2] * (double)s.esum) -
                                        s.hsum) - (double)L * (c.k[HBK_LNSQRT2PI] + s.lsd);
        if (!VIT && !a.gamma) return; // HB_WANT_LL alone: the forward sweep is the whole job
    }

    // ------------------------------------------------------------ backward sweep
    s.ycur = 0;
    if (VIT) {
        if (s.n0 == -INFINITY || s.n1 == -INFINITY || s.n2 == -INFINITY) *a.status = 1;
        double mv = s.n0;
        if (s.n1 > mv) { mv = s.n1; s.ycur = 1; }
        if (s.n2 > mv) { mv = s.n2; s.ycur = 2; }
    }
    s.b0 = s.b1 = s.b2 = 1.0;
    s.eb = 0;
    // stage the last full chunk while the tail (if any) is processed; second half first, because
    // the replay walks j = 0..7 but the refill order only has to respect the group accounting
    if (FB && nfull > 0) {
        const double *xl = xp + (int64_t)(nfull - 1) * HB_CHUNK * ld;
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) xs_refill(m, j, xl + (int64_t)j * ld);
            xs_commit(m);
        }
    }
    // The chunk's checkpoint and back-pointer word are fetched one chunk ahead (their latency
    // would otherwise sit on the critical path at every chunk boundary).
    uint64_t pk_nx = 0;
    double c0_nx = 0, c1_nx = 0, c2_nx = 0;
    if (nfull > 0) {
        if (VIT) pk_nx = psi[(int64_t)(nfull - 1) * sB];
        if (FB && nfull > 1) {
            const double *ck = ckpt + (int64_t)(nfull - 1) * 3 * sB;
            c0_nx = ck[0];
            c1_nx = ck[sB];
            c2_nx = ck[2 * sB];
        }
    }
    if (tail > 0) {
        uint64_t pk = 0;
        if (VIT) pk = psi[(int64_t)nfull * sB];
        if (FB) {
            if (nfull > 0) {
                const double *ck = ckpt + (int64_t)nfull * 3 * sB;
                s.f0 = ck[0];
                s.f1 = ck[sB];
                s.f2 = ck[2 * sB];
            }
            for (int j = 0; j < tail; j++) {
                const double xv = xp[(int64_t)(nfull * HB_CHUNK + j) * ld];
                norm3_replay_one<SYM>(c, s, xv, nfull == 0 && j == 0, m, j);
            }
        }
        for (int j = tail - 1; j >= 0; j--)
            norm3_backward_one<SYM, VIT, FB>(a, s, t0 + nfull * HB_CHUNK + j, b, pk, j, m);
    }
    for (int32_t ch = nfull - 1; ch >= 0; ch--) {
        const uint64_t pk = pk_nx;
        if (FB && ch > 0) {
            s.f0 = c0_nx;
            s.f1 = c1_nx;
            s.f2 = c2_nx;
        }
        if (ch > 0) {
            if (VIT) pk_nx = psi[(int64_t)(ch - 1) * sB];
            if (FB && ch > 1) {
                const double *ck = ckpt + (int64_t)(ch - 1) * 3 * sB;
                c0_nx = ck[0];
                c1_nx = ck[sB];
                c2_nx = ck[2 * sB];
            }
        }
        if (FB) {
            const double *xprev = xp + (int64_t)(ch - 1) * HB_CHUNK * ld;
            const bool more = ch > 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                xs_wait<1>(m);
                HB_UNROLL_STEPS
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    const double xv = m.xs[j * m.st];
                    if (more) xs_refill(m, j, xprev + (int64_t)j * ld);
                    norm3_replay_one<SYM>(c, s, xv, j == 0 && ch == 0, m, j);
                }
                xs_commit(m);
            }
        }
        const int64_t row0 = t0 + (int64_t)ch * HB_CHUNK;
        HB_UNROLL_STEPS
        for (int j = HB_CHUNK - 1; j >= 0; j--)
            norm3_backward_one<SYM, VIT, FB>(a, s, row0 + j, b, pk, j, m);
    }
    xs_wait<0>(m);
}

} // namespace hb
