// hb_norm3_fast.cuh — 3-state Gaussian chain, symmetric parameters: speculate-and-verify Viterbi
// fused with a normalisation-free forward-backward.  One thread per chain, same two-sweep layout
// as hb_norm3.cuh (which stays the exact / general-parameter path and the fallback of this one).
//
// Replaces HiddenMarkov::dthmm(..., "norm", ...) -> HiddenMarkov::Viterbi at
// R/HoneyBADGER.R:1150-1151 and R/HoneyBADGER_gexp.R:503-504 for the parameter block the reference
// always builds (R/HoneyBADGER.R:1144-1150: Pi = [[1-2t,t,t],[t,1-2t,t],[t,t,1-2t]], means
// (-dev, 0, +dev), one sd), and adds HiddenMarkov::forwardback's posteriors.
//
// Why a second formulation: hb_norm3.cuh replays the reference's arithmetic literally
// (3 correctly rounded divisions + dnorm per step, nu in absolute log scale): 122 fp64
// instructions per step, and the B200 fp64 pipe (2 warp-instructions/clk/SM) is what bounds the
// kernel, not HBM.  This file cuts the step to ~81 fp64 instructions:
//
//  Viterbi (12 DADD/step instead of ~40 fp64).  Work on DIFFERENCES d_k = nu_k - nu_1, normalised by
//  the "stay" cost: with tau = log t - log(1-2t) the three maxima become
//        m0 = max(d0, tau, d2+tau)   m1 = max(d0+tau, 0, d2+tau)   m2 = max(d0+tau, tau, d2)
//  (candidates in the reference's order j = 0,1,2; first index wins ties = strict ">" for a later j)
//  and d_k' = (m_k - m1) + (log b_k - log b_1),  log b_{0,2} - log b_1 = lnK -+ u,  u = (m/sd^2)(x-mu_1).
//  Every comparison is made on the SIGN of a rounded difference, whose magnitude is the decision
//  margin; the chain tracks the minimum margin (integer min on the high words, ALU pipe).  The
//  differences equal the reference's (computed with nu ~ 1e4 in absolute scale) up to an a-priori
//  bound on its accumulated rounding, E <= 2 u L M (u = 2^-53, M >= max |nu|, bounded from sum x^2):
//  if min margin > eps = 3 E every comparison provably has the reference's outcome, hence the same
//  back-pointers, final state and path — bit-exact.  Otherwise (a few chains in 10^5 for
//  per-chromosome chains) THE SAME THREAD re-runs the forward recursion with the
//  literal reference arithmetic of hb_norm3.cuh before the backward sweep.  Exact ties have margin 0
//  and always take that route, so "first index wins" is decided by the reference's own roundings.
//
//  Forward-backward (no per-step normalisation).  The filter phi_t is rescaled by 2^-es_t
//  (es_t = exponent of sum phi_{t-1}); the backward vector uses THE SAME es_t, which makes
//  sum_k phi_t(k) beta_t(k) independent of t.  Starting beta_T at 1/sum(phi_T) therefore yields
//  gamma_t(k) = phi_t(k) * beta_t(k) directly: 3 multiplies, no sum, no reciprocal.
//
// Layout, staging (cp.async ring for x), checkpoints every HB_CHUNK steps, back-pointer words and
// chunk-ahead prefetch are as in hb_norm3.cuh.  Build with -fmad=false (explicit dfma only).
#pragma once
#include "hb_norm3.cuh"

namespace hb {

// Longest chain the speculative path accepts: eps grows ~L^2, the flagged fraction ~L^3.
#define HB_FAST_LMAX 4096
#ifndef HB_FAST_PF_CHUNKS
#define HB_FAST_PF_CHUNKS 2 // L2 prefetch distance (chunks) ahead of the staging of x / checkpoints
// (Tried and rejected: 2-double checkpoints — state divided by its largest component, position of that
// component in the free sign bits, beta renormalised at every chunk entry.  8 B less scratch per chunk
// each way (48.4 -> 46.4 B/step) but one more division per chunk and sweep and 6 more registers:
// FB-only 1.549 -> 1.521 ms, Viterbi + FB 1.940 -> 1.954 ms; again after the instruction-count work
// (cooperative staging etc.): FB-only 1.55 -> 1.51 ms, Viterbi + FB 1.85 -> 1.88 ms.  Not kept.)
#endif

// ---------------------------------------------------------------- small bit helpers
// (code << 1) | signbit(q)
HB_HD uint32_t push_sign(uint32_t code, double q)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)hi32(q), code, 1);
#else
    return (code << 1) | ((uint32_t)hi32(q) >> 31);
#endif
}
// running minimum of |q| on the high word (monotone in |q|; NaN/Inf map above every finite value)
HB_HD uint32_t absmin_hi(uint32_t mn, double q)
{
    const uint32_t h = (uint32_t)hi32(q) & 0x7fffffffu;
    return h < mn ? h : mn;
}
// q < 0 (by sign bit) ? a : b
HB_HD double sel_neg(double q, double a, double b) { return hi32(q) < 0 ? a : b; }

// exp(u), exp(-u) as Pp * 2^n, Pm * 2^-n: one-FMA range reduction (|n| small on this path: the
// reduction error is n * 2.3e-17), even/odd split of the degree-11 Taylor polynomial on
// |r| <= ln2/2 (truncation 6e-15 at the interval ends, 5e-16 on average).  17 fp64 instructions.
HB_HD void fast_exp_pair(const Norm3Const &c, double u, double &Pp, double &Pm, int32_t &n)
{
    const double MAGIC = 6755399441055744.0; // 1.5 * 2^52
    const double t = dfma(u, c.k[HBK_L2E], MAGIC);
    const int32_t ni = lo32(t);
    const double nf = t - MAGIC;
    const double r = dfma(nf, -c.k[HBK_LN2], u);
    const double q = r * r;
    double ev = dfma(q, c.k[HBK_E10], c.k[HBK_E8]);
    double od = dfma(q, c.k[HBK_O11], c.k[HBK_O9]);
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

// Per-chain state of both sweeps.
struct Fast3State {
    double A, Bc, lnK;    // u = A x + Bc ; log(b_{0,2}/b_1) = lnK -+ u
    double K, aK, tK;     // filter: rho_{0,2} = K e^{-+u}; aK = K (a - t), tK = K t
    double d0, d2;        // Viterbi differences
    double sx2;           // sum x^2 (error bound), or sum (x - mu_1)^2 when the log-likelihood is wanted
    int32_t esum;         // sum of removed binary exponents (log-likelihood)
    uint32_t mn;          // min |decision difference|, high word
    double f0, f1, f2;    // forward filter
    double b0, b1, b2;    // backward vector (pre-divided by sum phi_T)
    int32_t ycur;
};

// One Viterbi step on differences; returns the 6-bit code (state k's field at bit 4-2k holds
// b_k<<1 | a_k: a_k = "j=1 beat j=0", b_k = "j=2 beat the winner"; back-pointer = min(field, 2)).
HB_HD uint32_t fast_vit_step(const Norm3Const &c, Fast3State &s, double u, bool first)
{
    const double tau = c.tau;
    const double e0 = s.lnK - u, e2 = s.lnK + u;
    if (first) {
        s.d0 = c.dl0 + e0;
        s.d2 = c.dl2 + e2;
        return 0;
    }
    const double d0 = s.d0, d2 = s.d2;
    const double C0 = d0 + tau, C2 = d2 + tau;
    uint32_t code = 0, mn = s.mn;
    // state 0: d0 | tau | C2
    const double qa0 = d0 - tau;
    const double w0 = sel_neg(qa0, tau, d0);
    const double qb0 = w0 - C2;
    const double m0 = sel_neg(qb0, C2, w0);
    code = push_sign(push_sign(code, qb0), qa0);
    mn = absmin_hi(absmin_hi(mn, qa0), qb0);
    // state 1: C0 | 0 | C2          (0 > C0  <=>  C0 < 0)
    const double w1 = sel_neg(C0, 0.0, C0);
    const double qb1 = w1 - C2;
    const double m1 = sel_neg(qb1, C2, w1);
    code = push_sign(push_sign(code, qb1), C0);
    mn = absmin_hi(absmin_hi(mn, C0), qb1);
    // state 2: C0 | tau | d2        (tau > C0  <=>  nu_1 > nu_0  <=>  d0 < 0)
    const double w2 = sel_neg(d0, tau, C0);
    const double qb2 = w2 - d2;
    const double m2 = sel_neg(qb2, d2, w2);
    code = push_sign(push_sign(code, qb2), d0);
    mn = absmin_hi(absmin_hi(mn, d0), qb2);
    s.mn = mn;
    s.d0 = (m0 - m1) + e0;
    s.d2 = (m2 - m1) + e2;
    return code;
}

// Filter step phi <- ((phi Pi) . rho) 2^-es; returns the scaled ratios and the high word of 2^-es
// (what the backward sweep needs).  `first`: phi = delta . rho, es = 0.
HB_HD int32_t fast_fwd_step(const Norm3Const &c, Fast3State &s, double u, bool first, double &E0s,
                            double &E2s, int32_t &ph)
{
    double Pp, Pm;
    int32_t n;
    fast_exp_pair(c, u, Pp, Pm, n);
    double p0, p1, p2;
    int32_t es;
    if (first) {
        p0 = c.delta[0] * s.K;
        p1 = c.delta[1];
        p2 = c.delta[2] * s.K;
        es = 0;
    } else {
        const double sm = (s.f0 + s.f1) + s.f2;
        es = expo(sm);
        const double ts = c.tt * sm, tKs = s.tK * sm;
        p1 = dfma(c.amt, s.f1, ts);
        p0 = dfma(s.aK, s.f0, tKs);
        p2 = dfma(s.aK, s.f2, tKs);
    }
    ph = (1023 - es) << 20;
    E0s = scale2(Pm, -n - es);
    E2s = scale2(Pp, n - es);
    s.f1 = p1 * mk64(ph, 0);
    s.f0 = p0 * E0s;
    s.f2 = p2 * E2s;
    return es;
}

// ---------------------------------------------------------------- shared-memory slots
// Compile-time layout switches (tools/build_variant.sh explores them):
//   HB_FAST_NOPHI1 1: phi_1 is not parked, gamma_1 = |1 - gamma_0 - gamma_2| (the posteriors of this
//                  formulation sum to 1 within ~1e-13 by construction)
#ifndef HB_FAST_NOPHI1
#define HB_FAST_NOPHI1 1
#endif
// per step j and thread: va, vb (double2), [e2 (double)], ph = hi word of 2^-es, [xs = x ring]
//   NOPHI1: va = (phi0, phi2), vb = (E0s, E2s);  else va = (phi0, phi1), vb = (phi2, E0s), e2 = E2s
#define HB_FAST_PARK_BYTES (16 + 16 + (HB_FAST_NOPHI1 ? 0 : 8) + 4)
#define HB_FAST_SMEM_BYTES_PER_THREAD (HB_CHUNK * (HB_FAST_PARK_BYTES + 8 + 1))
// bulk (TMA) staging on the tiled layout: two chunks of x per warp + two mbarriers per warp
#define HB_FAST_BULK_SMEM_BYTES(nthreads) \
    ((size_t)HB_CHUNK * HB_FAST_PARK_BYTES * (nthreads) + (size_t)2 * HB_CHUNK * 8 * (nthreads) + 16 * ((nthreads) / 32))

HB_HD void fast_replay_one(const Norm3Const &c, Fast3State &s, double xv, bool first, Norm3Smem &m,
                           int j)
{
    double E0s, E2s;
    int32_t ph;
    fast_fwd_step(c, s, dfma(xv, s.A, s.Bc), first, E0s, E2s, ph);
    double2 a, b;
#if HB_FAST_NOPHI1
    a.x = s.f0;
    a.y = s.f2;
    b.x = E0s;
    b.y = E2s;
#else
    a.x = s.f0;
    a.y = s.f1;
    b.x = s.f2;
    b.y = E0s;
    m.e2[j * m.st] = E2s;
#endif
    m.va[j * m.st] = a;
    m.vb[j * m.st] = b;
    m.ph[j * m.st] = ph;
}

// Backward step at one row: emit y and gamma, move beta one position back.  pw = the 32-bit word of
// the chunk's back-pointer word that holds step j, sh = bit offset of step j's code inside it.
// g0/g1/g2 point at this row in the three posterior planes.
// YS: the state goes to the warp's shared-memory tile (m.ys) instead of global memory; the caller
// writes the tile out once per chunk.
template <bool VIT, bool FB, bool YS = false>
HB_HD void fast_backward_one(const Norm3Const &c, Fast3State &s, uint8_t *yp, double *g0, double *g1,
                             double *g2, uint32_t pw, int sh, int j, const Norm3Smem &m)
{
    if (VIT) {
        if (YS)
            m.ys[j * 32] = (uint8_t)(s.ycur + 1);
        else
            st_stream(yp, (uint8_t)(s.ycur + 1));
        const uint32_t f = (pw >> (sh + 4 - 2 * s.ycur)) & 3u;
        s.ycur = (int32_t)(f < 2u ? f : 2u);
    }
    if (FB) {
        const double2 va = m.va[j * m.st], vb = m.vb[j * m.st];
        const double p2 = mk64(m.ph[j * m.st], 0);
#if HB_FAST_NOPHI1
        const double ga = va.x * s.b0, gc = va.y * s.b2;
        st_stream(g0, ga);
        st_stream(g2, gc);
        st_stream(g1, fabs(1.0 - (ga + gc)));
        const double v0 = vb.x * s.b0, v2 = vb.y * s.b2, v1 = p2 * s.b1;
#else
        const double E2s = m.e2[j * m.st];
        *g0 = va.x * s.b0;
        *g1 = va.y * s.b1;
        *g2 = vb.x * s.b2;
        const double v0 = vb.y * s.b0, v2 = E2s * s.b2, v1 = p2 * s.b1;
#endif
        const double sw = dfma(s.K, v0 + v2, v1);
        const double tsw = c.tt * sw;
        s.b1 = dfma(c.amt, v1, tsw);
        s.b0 = dfma(s.aK, v0, tsw);
        s.b2 = dfma(s.aK, v2, tsw);
    }
}

// Warp-cooperative staging on the tiled layout: a half chunk of a warp's x is HB_HALF rows of 256
// contiguous bytes, copied as 16-byte pieces — two cp.async per lane and half instead of one 8-byte
// cp.async per lane and row.  g = this chain's element of the half's first row, lane_c = this chain's
// column inside its tile (pointer bases), lane = the executing lane (which pieces it moves).  Lanes
// read slots that other lanes filled: a __syncwarp follows every wait and precedes every refill.
// LAST: the backward sweep's reads are the last use of x — with -DHB_BWD_EVICT_FIRST they carry an L2
// evict-first policy so that the forward sweeps' freshly read tails survive longer (experiment).
template <bool LAST = false>
HB_HD void xs_refill_half_coop(Norm3Smem &m, int h, const double *g, int lane_c, int lane)
{
#if defined(__CUDA_ARCH__)
    const double *gb = g - lane_c;
    double *sb = m.xs - lane_c + h * HB_HALF * m.st;
#if defined(HB_BWD_EVICT_FIRST)
    uint64_t pol = 0;
    if (LAST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
#pragma unroll
    for (int q = 0; q < HB_HALF / 2; q++) {
        const int p = lane + 32 * q, r = p >> 4, e = (p & 15) * 2;
        unsigned sa = (unsigned)__cvta_generic_to_shared(sb + r * m.st + e);
#if defined(HB_BWD_EVICT_FIRST)
        if (LAST) {
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa),
                         "l"(gb + r * 32 + e), "l"(pol)
                         : "memory");
            continue;
        }
#endif
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gb + r * 32 + e) : "memory");
    }
#else
    (void)m, (void)h, (void)g, (void)lane_c, (void)lane;
#endif
}
HB_HD void warp_sync()
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

// bit position of step j's 6-bit code: 5 steps per 32-bit word
#define HB_PSI_WORD(j) ((j) / 5)
#define HB_PSI_SHIFT(j) (6 * ((j) % 5))

// ---------------------------------------------------------------- exact fallback (rare)
// Recomputes the chain's back-pointer words with the literal reference arithmetic (hb_norm3.cuh),
// reading x straight from global memory; returns the final state.  Same word format as the fast path.
HB_HD_NOINLINE int32_t norm3_exact_forward(const Norm3Args &a, int32_t seg, int64_t b)
{
    const Norm3Const &c = a.c;
    const int64_t t0 = a.seg_off[seg];
    const int32_t L = (int32_t)(a.seg_off[seg + 1] - t0);
    const int64_t pix = a.sd_per_seg ? (int64_t)seg * a.B + b : b;
    const double sdv = a.sd[pix], isd = a.inv_sd[pix], lsd = a.log_sd[pix];
    const int64_t ld = lay_stride(a.tiled, a.ld), sB = lay_stride(a.tiled, a.B);
    const double *xp = a.x + lay_off(a.tiled, a.T, a.ld, t0, b);
    uint64_t *psi = a.psi + lay_off(a.tiled, a.nchunks, a.B, a.chunk_off[seg], b);
    const int32_t nch = (L + HB_CHUNK - 1) / HB_CHUNK;
    double n0 = 0, n1 = 0, n2 = 0;
    // Only this lane of the warp is active: keep one chunk of x in flight ahead of the recursion so
    // the DRAM latency of the strided loads overlaps the (latency-bound) exact arithmetic.
    double xc[HB_CHUNK], xn[HB_CHUNK];
#pragma unroll
    for (int j = 0; j < HB_CHUNK; j++) xc[j] = j < L ? xp[(int64_t)j * ld] : 0.0;
    for (int32_t ch = 0; ch < nch; ch++) {
        const int32_t i0 = ch * HB_CHUNK;
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++)
            xn[j] = (i0 + HB_CHUNK + j < L) ? xp[(int64_t)(i0 + HB_CHUNK + j) * ld] : 0.0;
        uint32_t w0 = 0u, w1 = 0u;
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) {
            if (i0 + j < L) {
                double l0, l1, l2, h1;
                norm3_logb(xc[j], c, sdv, isd, lsd, l0, l1, l2, h1);
                if (i0 + j == 0) {
                    n0 = c.logdelta[0] + l0;
                    n1 = c.logdelta[1] + l1;
                    n2 = c.logdelta[2] + l2;
                } else {
                    const uint32_t ps = norm3_vit_step<true>(c, n0, n1, n2, l0, l1, l2);
                    const uint32_t code = ((ps & 3u) << 4) | (ps & 12u) | ((ps >> 4) & 3u);
                    if (HB_PSI_WORD(j))
                        w1 |= code << HB_PSI_SHIFT(j);
                    else
                        w0 |= code << HB_PSI_SHIFT(j);
                }
            }
        }
        psi[(int64_t)ch * sB] = (uint64_t)w0 | ((uint64_t)w1 << 32);
#pragma unroll
        for (int j = 0; j < HB_CHUNK; j++) xc[j] = xn[j];
    }
    if (n0 == -INFINITY || n1 == -INFINITY || n2 == -INFINITY) *a.status = 1;
    int32_t y = 0;
    double mv = n0;
    if (n1 > mv) { mv = n1; y = 1; }
    if (n2 > mv) { mv = n2; y = 2; }
    return y;
}

HB_HD void count_flagged(const Norm3Args &a, int32_t seg, int64_t b)
{
#if defined(__CUDA_ARCH__)
    const int32_t i = atomicAdd(a.status + 1, 1);
#else
    const int32_t i = a.status[1]++;
#endif
    if (a.flagged && i < a.flagged_cap) a.flagged[i] = (int64_t)seg * a.B + b;
}

// ---------------------------------------------------------------- the whole chain
// sum x^2 (LL = false) or sum (x - mu_1)^2 (LL = true: also the quadratic part of sum log b_1)
template <bool LL>
HB_HD void fast_acc_sq(const Norm3Const &c, Fast3State &s, double xv)
{
    if (LL) {
        const double dd = xv - c.mean[1];
        s.sx2 = dfma(dd, dd, s.sx2);
    } else {
        s.sx2 = dfma(xv, xv, s.sx2);
    }
}

// BULK (tiled layout only): x is staged by cp.async.bulk, one 2 KB copy per warp and chunk (BulkRing)
// TL: the layout is the warp-tiled one at compile time — every stride is the constant 32, so the
// row-to-row pointer updates of x, y, the three posterior planes and the scratch fold into immediate
// offsets (the gene-major form keeps run-time strides).
// live = false: this thread is a stand-in that repeats the last real chain of a partially filled tile
// so that the warp stays complete for the cooperative staging; it writes the same values to the same
// places and must only stay out of the counters.
template <bool VIT, bool FB, bool LL, bool BULK, bool TL = false>
HB_HD void norm3_fast_chain(const Norm3Args &a, int32_t seg, int64_t b, Norm3Smem &m, BulkRing &br,
                            bool live = true)
{
    const int32_t tiled = TL ? 1 : a.tiled;
#if defined(__CUDA_ARCH__)
    constexpr bool COOP = TL && !BULK;
    const int lane = (int)(threadIdx.x & 31u);
#else
    constexpr bool COOP = false;
    const int lane = 0;
#endif
    const int lane_c = (int)(b & 31);
    const Norm3Const &c = a.c;
    const int64_t t0 = a.seg_off[seg];
    const int32_t L = (int32_t)(a.seg_off[seg + 1] - t0);
    if (L <= 0) return;
    const int32_t nfull = L / HB_CHUNK;
    const int32_t tail = L - nfull * HB_CHUNK;
    const int64_t cbase = a.chunk_off[seg];
    const int64_t pix = a.sd_per_seg ? (int64_t)seg * a.B + b : b;
    const int64_t ld = lay_stride(tiled, a.ld), sB = lay_stride(tiled, a.B);
    const int64_t sy = lay_stride(tiled, a.ldy), sg = lay_stride(tiled, a.ldg);
    const double *xp = a.x + lay_off(tiled, a.T, a.ld, t0, b);
    uint64_t *psi = a.psi + lay_off(tiled, a.nchunks, a.B, cbase, b);
    double *ckpt = a.ckpt + (tiled ? (((b >> 5) * a.nchunks + cbase) * 3 << 5) + (b & 31) : cbase * 3 * a.B + b);
    uint8_t *ybase = a.y ? a.y + lay_off(tiled, a.T, a.ldy, t0, b) : nullptr;          // row t0
    double *gbase = a.gamma ? a.gamma + lay_off(tiled, a.T, a.ldg, t0, b) : nullptr; // row t0, plane 0

    Fast3State s;
    const double isd = a.inv_sd[pix];
    {
        const double mm = c.mean[2] - c.mean[1];
        const double mi = mm * isd;
        s.A = mi * isd;
        s.Bc = -s.A * c.mean[1];
        s.lnK = -0.5 * mi * mi;
        s.K = exp(s.lnK);
        s.aK = s.K * c.amt;
        s.tK = s.K * c.tt;
    }
    s.d0 = s.d2 = 0;
    s.sx2 = 0;
    s.esum = 0;
    s.mn = 0x7fffffffu;
    s.f0 = s.f1 = s.f2 = 0;
    const L2Pol pol = l2pol_make();

    // ------------------------------------------------------------ forward sweep
    // x reaches the arithmetic one chunk ahead of its use: cp.async ring, or bulk copies (BULK)
    if (BULK) {
        if (nfull > 0) bulk_issue<HB_CHUNK>(br, 0, xp);
        if (nfull > 1) bulk_issue<HB_CHUNK>(br, 1, xp + (int64_t)HB_CHUNK * ld);
    } else if (nfull > 0) {
        const double *xr0 = xp;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (COOP) {
                xs_refill_half_coop(m, h, xr0, lane_c, lane);
                xr0 += HB_HALF * ld;
            } else {
#pragma unroll
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    xs_refill(m, j, xr0);
                    xr0 += ld;
                }
            }
            xs_commit(m);
        }
    }
    {
        const double *xr = xp + (int64_t)HB_CHUNK * ld; // refill pointer: one chunk ahead
        uint64_t *pw = psi;
        double *ck = ckpt;
        for (int32_t ch = 0; ch < nfull; ch++) {
            const bool more = ch + 1 < nfull;
            const bool pf = ch + 1 + HB_FAST_PF_CHUNKS < nfull;
            (void)pf;
            uint32_t w[2] = {0u, 0u};
            const int stg = ch & 1;
            if (BULK) bulk_wait(br, stg);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (!BULK) xs_wait<1>(m);
                if (COOP) warp_sync();
                const double *xh = xr; // this chain's element of the next chunk's row h * HB_HALF
                HB_UNROLL_STEPS
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    double xv;
                    if (BULK) {
                        xv = bulk_read<HB_CHUNK>(br, stg, j);
                    } else {
                        xv = m.xs[j * m.st];
                        if (!COOP && more) xs_refill(m, j, xr);
                        if (TL) {
                            if (j == 0 && pf)
                                prefetch_l2_chunk(xr + (int64_t)HB_FAST_PF_CHUNKS * HB_CHUNK * 32, (int)(b & 31),
                                                  HB_CHUNK * 2);
                        } else if (pf) {
                            prefetch_l2(xr + (int64_t)HB_FAST_PF_CHUNKS * HB_CHUNK * ld);
                        }
                        xr += ld;
                    }
                    const bool first = (j == 0) && (ch == 0);
                    const double u = dfma(xv, s.A, s.Bc);
                    if (VIT || LL) fast_acc_sq<LL>(c, s, xv);
                    if (VIT) w[HB_PSI_WORD(j)] |= fast_vit_step(c, s, u, first) << HB_PSI_SHIFT(j);
                    if (FB) {
                        double E0s, E2s;
                        int32_t ph;
                        const int32_t es = fast_fwd_step(c, s, u, first, E0s, E2s, ph);
                        if (LL) s.esum += es;
                    }
                }
                if (COOP) {
                    warp_sync(); // every lane has read this half's rows
                    if (more) xs_refill_half_coop(m, h, xh, lane_c, lane);
                }
                if (!BULK) xs_commit(m);
            }
            if (BULK) {
                bulk_release(br);
                if (ch + 2 < nfull) bulk_issue<HB_CHUNK>(br, stg, xp + (int64_t)(ch + 2) * HB_CHUNK * ld);
            }
            if (VIT) st_keep(pol, pw, (uint64_t)w[0] | ((uint64_t)w[1] << 32));
            pw += sB;
            ck += 3 * sB;
            if (FB && (more || tail > 0)) {
                st_keep(pol, ck, s.f0);
                st_keep(pol, ck + sB, s.f1);
                st_keep(pol, ck + 2 * sB, s.f2);
            }
        }
    }
    if (!BULK) xs_wait<0>(m);
    if (tail > 0) { // rolled: at most HB_CHUNK-1 steps per chain
        uint32_t w[2] = {0u, 0u};
        for (int j = 0; j < tail; j++) {
            const double xv = xp[(int64_t)(nfull * HB_CHUNK + j) * ld];
            const bool first = nfull == 0 && j == 0;
            const double u = dfma(xv, s.A, s.Bc);
            if (VIT || LL) fast_acc_sq<LL>(c, s, xv);
            if (VIT) {
                const uint32_t code = fast_vit_step(c, s, u, first);
                if (j < 5)
                    w[0] |= code << (6 * j);
                else
                    w[1] |= code << (6 * (j - 5));
            }
            if (FB) {
                double E0s, E2s;
                int32_t ph;
                const int32_t es = fast_fwd_step(c, s, u, first, E0s, E2s, ph);
                if (LL) s.esum += es;
            }
        }
        if (VIT) psi[(int64_t)nfull * sB] = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    }

    if (FB && LL) {
        // LL = log sum(alpha_L) = log(sum phi) + ln2 * esum + sum_t log b_1(x_t)
        const double lsd = a.log_sd[pix];
        a.ll[(int64_t)seg * a.B + b] =
            ((log((s.f0 + s.f1) + s.f2) + c.k[HBK_LN2] * (double)s.esum) - 0.5 * (s.sx2 * (isd * isd))) -
            (double)L * (c.k[HBK_LNSQRT2PI] + lsd);
        if (!VIT && !a.gamma) return; // HB_WANT_LL alone: the forward sweep is the whole job
    }
    // ------------------------------------------------------------ verify the speculation
    s.ycur = 0;
    if (VIT) {
        // final state: nu_1 > nu_0 <=> d0 < 0 ; nu_2 > max <=> d2 > max(d0, 0)
        const double wf = sel_neg(s.d0, 0.0, s.d0);
        const double qf = wf - s.d2;
        s.ycur = hi32(s.d0) < 0 ? 1 : 0;
        if (hi32(qf) < 0) s.ycur = 2;
        uint32_t mn = absmin_hi(absmin_hi(s.mn, s.d0), qf);
        // a-priori bound on |reference difference - ours| (see the header): u (2 L M + 16 M + 16 L dmax)
        const double U = 1.1102230246251565e-16;
        const double Ld = (double)L;
        const double lsd = a.log_sd[pix];
        // |log b_k| <= C + |log sd| + ((x - r)^2 + (mu_k - r)^2) / sd^2 with r = 0 (sx2 = sum x^2) or
        // r = mu_1 (LL); |u| <= |A| sqrt(sx2) + |A r - A mu_1|
        const double M = Ld * ((c.k[HBK_LNSQRT2PI] + fabs(lsd)) + fabs(c.la)) +
                         (s.sx2 + Ld * (LL ? c.dmu2 : c.mumax2)) * (isd * isd) + (fabs(c.lt) + c.ldabs);
        const double dmax = ((fabs(c.tau) + fabs(s.lnK)) + fabs(s.A) * sqrt(s.sx2)) +
                            ((LL ? 0.0 : fabs(s.Bc)) + 8.0);
        const double eps = 3.0 * U * ((2.0 * Ld * M + 16.0 * M) + 16.0 * Ld * dmax) * c.eps_scale;
        const double margin = mk64((int32_t)mn, 0); // <= the true minimum |difference|
        if (!(margin > eps)) {
            if (live) count_flagged(a, seg, b);
            s.ycur = norm3_exact_forward(a, seg, b);
        } else if (s.d0 == -INFINITY || s.d2 == -INFINITY) {
            *a.status = 1; // some nu_k(T) == -Inf (nu_1 is finite here): HiddenMarkov's underflow stop
        }
    }

    // ------------------------------------------------------------ backward sweep
    if (FB) {
        const double q = 1.0 / ((s.f0 + s.f1) + s.f2);
        s.b0 = s.b1 = s.b2 = q;
    }
    if (BULK) {
        if (FB && nfull > 0) bulk_issue<HB_CHUNK>(br, 0, xp + (int64_t)(nfull - 1) * HB_CHUNK * ld);
        if (FB && nfull > 1) bulk_issue<HB_CHUNK>(br, 1, xp + (int64_t)(nfull - 2) * HB_CHUNK * ld);
    } else if (FB && nfull > 0) {
        const double *xl = xp + (int64_t)(nfull - 1) * HB_CHUNK * ld;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (COOP) {
                xs_refill_half_coop<true>(m, h, xl, lane_c, lane);
                xl += HB_HALF * ld;
            } else {
#pragma unroll
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    xs_refill(m, j, xl);
                    xl += ld;
                }
            }
            xs_commit(m);
        }
    }
    uint64_t pk_nx = 0;
    double c0_nx = 0, c1_nx = 0, c2_nx = 0;
    if (nfull > 0) {
        if (VIT) pk_nx = ld_drop(pol, psi + (int64_t)(nfull - 1) * sB);
        if (FB && nfull > 1) {
            const double *ck = ckpt + (int64_t)(nfull - 1) * 3 * sB;
            c0_nx = ld_drop(pol, ck);
            c1_nx = ld_drop(pol, ck + sB);
            c2_nx = ld_drop(pol, ck + 2 * sB);
        }
    }
    const int64_t plane = a.gplane;
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
                fast_replay_one(c, s, xv, nfull == 0 && j == 0, m, j);
            }
        }
        for (int j = tail - 1; j >= 0; j--) {
            const int64_t rel = nfull * HB_CHUNK + j; // position within the chain
            const uint32_t pwd = j < 5 ? (uint32_t)pk : (uint32_t)(pk >> 32);
            double *gr = FB ? gbase + rel * sg : nullptr;
            fast_backward_one<VIT, FB>(c, s, VIT ? ybase + rel * sy : nullptr, gr, gr + plane,
                                       gr + 2 * plane, pwd, j < 5 ? 6 * j : 6 * (j - 5), j, m);
        }
    }
    {
        // row pointers of the LAST row of the current chunk, stepped back one row per step
        const int64_t rell = (int64_t)nfull * HB_CHUNK - 1;
        uint8_t *yp = VIT ? ybase + rell * sy : nullptr;
        double *g0 = FB ? gbase + rell * sg : nullptr;
        double *g1 = g0 + plane, *g2 = g1 + plane;
        const double *xr = xp + (int64_t)(nfull - 2) * HB_CHUNK * ld; // refill: previous chunk
        const uint64_t *pr = psi + (int64_t)(nfull - 2) * sB;
        const double *cr = ckpt + (int64_t)(nfull - 2) * 3 * sB;
        for (int32_t ch = nfull - 1; ch >= 0; ch--) {
            const uint64_t pk = pk_nx;
            if (FB && ch > 0) {
                s.f0 = c0_nx;
                s.f1 = c1_nx;
                s.f2 = c2_nx;
            }
            if (ch > 0) {
                if (VIT) pk_nx = ld_drop(pol, pr);
                if (FB && ch > 1) {
                    c0_nx = ld_drop(pol, cr);
                    c1_nx = ld_drop(pol, cr + sB);
                    c2_nx = ld_drop(pol, cr + 2 * sB);
                }
            }
            const bool pf = ch > 1 + HB_FAST_PF_CHUNKS;
            if (pf) { // checkpoints / back-pointer words of the chunk HB_FAST_PF_CHUNKS further back
                if (VIT) prefetch_l2(pr - (int64_t)HB_FAST_PF_CHUNKS * sB);
                if (FB) {
                    const double *cf = cr - (int64_t)HB_FAST_PF_CHUNKS * 3 * sB;
                    prefetch_l2(cf);
                    prefetch_l2(cf + sB);
                    prefetch_l2(cf + 2 * sB);
                }
            }
            pr -= sB;
            cr -= 3 * sB;
            if (FB) {
                const bool more = ch > 0;
                const double *xq = xr;
                const int stg = (nfull - 1 - ch) & 1;
                if (BULK) bulk_wait(br, stg);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (!BULK) xs_wait<1>(m);
                    if (COOP) warp_sync();
                    const double *xh = xq;
                    HB_UNROLL_STEPS
                    for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                        double xv;
                        if (BULK) {
                            xv = bulk_read<HB_CHUNK>(br, stg, j);
                        } else {
                            xv = m.xs[j * m.st];
                            if (!COOP && more) xs_refill(m, j, xq);
                            if (TL) {
                                if (j == 0 && pf)
                                    prefetch_l2_chunk(xq - (int64_t)HB_FAST_PF_CHUNKS * HB_CHUNK * 32,
                                                      (int)(b & 31), HB_CHUNK * 2);
                            } else if (pf) {
                                prefetch_l2(xq - (int64_t)HB_FAST_PF_CHUNKS * HB_CHUNK * ld);
                            }
                            xq += ld;
                        }
                        fast_replay_one(c, s, xv, j == 0 && ch == 0, m, j);
                    }
                    if (COOP) {
                        warp_sync();
                        if (more) xs_refill_half_coop<true>(m, h, xh, lane_c, lane);
                    }
                    if (!BULK) xs_commit(m);
                }
                if (BULK) {
                    bulk_release(br);
                    if (ch >= 2) bulk_issue<HB_CHUNK>(br, stg, xp + (int64_t)(ch - 2) * HB_CHUNK * ld);
                }
                xr -= (int64_t)HB_CHUNK * ld;
            }
            const uint32_t plo = (uint32_t)pk, phi = (uint32_t)(pk >> 32);
#ifdef HB_FAST_BWD_UNROLL
#pragma unroll HB_FAST_BWD_UNROLL
#else
#pragma unroll
#endif
            for (int j = HB_CHUNK - 1; j >= 0; j--) {
                fast_backward_one<VIT, FB, COOP>(c, s, yp, g0, g1, g2, HB_PSI_WORD(j) ? phi : plo,
                                                 HB_PSI_SHIFT(j), j, m);
                if (VIT && !COOP) yp -= sy;
#ifndef HB_DEBUG_NOGAMMA // (debug variant: keep writing the same rows -> the stores stay in L2)
                if (FB) {
                    g0 -= sg;
                    g1 -= sg;
                    g2 -= sg;
                }
#endif
            }
            if (COOP && VIT) {
                // the chunk's states: 8 rows x 32 bytes of the warp = 256 contiguous bytes, 8 per lane
                warp_sync();
                const uint64_t v = *reinterpret_cast<const uint64_t *>(m.ys - lane_c + lane * 8);
                st_stream(reinterpret_cast<uint64_t *>(ybase + (int64_t)ch * HB_CHUNK * sy - lane_c + lane * 8), v);
                if (!FB) warp_sync(); // (with FB the next chunk's staging syncs order the tile's reuse)
            }
        }
    }
    if (!BULK) xs_wait<0>(m);
}

} // namespace hb
