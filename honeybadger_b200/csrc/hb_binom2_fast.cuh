// hb_binom2_fast.cuh — 2-state binomial chain, lean path: literal Viterbi + normalisation-free
// forward-backward in plain doubles.  One thread per chain; hb_binom2.cuh stays the general path
// (relative-exponent arithmetic for evidence beyond 2^+-900 per position) and re-runs the launch
// when this kernel raises its gate.
//
// Replaces HiddenMarkov::dthmm(mafl, Pi, delta, "binom", list(prob=c(pd,pn)), list(size=sizel),
// discrete=TRUE) -> HiddenMarkov::Viterbi at R/HoneyBADGER.R:1409-1410 and
// R/HoneyBADGER_allele.R:579-580, plus HiddenMarkov::forwardback's posteriors.
//
// Viterbi: the reference's operation order on host-built log dbinom values (table lookup, see
// hb_binom2.cuh) — per-cell allele chains tie EXACTLY at most zero-coverage positions
// ((a + log t) + log(1-t) vs (a + log(1-t)) + log t), so only the literal arithmetic decides
// "first index wins" the way R does; there is nothing to speculate on here.
//
// Forward-backward: phi_t = ((phi_{t-1} Pi) . (rho_t, 1)) 2^-es_t with es_t the exponent of the
// un-normalised sum (phi stays in [1, 2)); the backward vector applies the SAME 2^-es_t, so
// sum_k phi_t(k) beta_t(k) is constant and, with beta_T = 1 / sum(phi_T),
//        gamma_t(0) = phi_t(0) beta_t(0),   gamma_t(1) = |1 - gamma_t(0)|.
// rho = b_0 / b_1 comes from a table of doubles with the exponent limited to +-900 (entries beyond
// that are marked and raise the gate).  Since rows of Pi are positive, beta's two components stay
// within max Pi_jk / Pi_j'k of each other and beta <= const / max phi: nothing can overflow.
//
// 32 fp64 instructions per step (Viterbi 8, filter 8, replay 8, backward 8) against ~90 fp64 +
// ~280 integer instructions of the general path.  HBM per step: x, n twice (16 B), gamma 16, y 1,
// back-pointers 2 bits (0.5 B w+r), checkpoints 2 doubles / 8 steps (4 B w+r) = 37.5 B (25 B algorithmic).
#pragma once
#include "hb_binom2.cuh"

namespace hb {

#define HB_B2F_RHO_EXP_MAX 900
#ifndef HB_B2F_PF_CHUNKS
#define HB_B2F_PF_CHUNKS 2 // L2 prefetch distance of the count planes, in chunks
#endif

// per step and thread: (phi0, rho 2^-es) as double2 + high word of 2^-es + the two count rings
#define HB_B2F_SMEM_BYTES_PER_THREAD (HB_CHUNK * (16 + 4 + 4 + 4))

// Emission table entries with n <= HB_B2F_SLUT_N live in shared memory (one copy per block): the
// per-step lookup is then an LDS with a ~30-cycle latency instead of an L1/L2 gather that the
// streaming traffic keeps evicting (ncu: 73 % of the stall samples were long-scoreboard waits).
#define HB_B2F_SLUT_N 31
#define HB_B2F_SLUT_ENTRIES ((HB_B2F_SLUT_N + 1) * (HB_B2F_SLUT_N + 2) / 2)
#define HB_B2F_SLUT_BYTES (HB_B2F_SLUT_ENTRIES * 24)

// back-pointer word of one chunk: 2 bits / step
#if HB_CHUNK <= 8
typedef uint16_t b2f_psi_t;
#else
typedef uint32_t b2f_psi_t;
#endif

struct B2fSmem {
    double2 *pr;
    int32_t *ph;
    int32_t *xs, *ns; // cp.async rings of the count planes, one chunk deep
    int32_t st;
    const double2 *slb; // [HB_B2F_SLUT_ENTRIES] (l0, l1)
    const double *srho; // [HB_B2F_SLUT_ENTRIES]
    int32_t slut_n;     // min(HB_B2F_SLUT_N, lut_nmax)
};

// rho as a double for the table / the slow path: negative = "beyond the representable range"
HB_HD double b2f_rho_double(const RhoME &r)
{
    if (r.e > HB_B2F_RHO_EXP_MAX || r.e < -HB_B2F_RHO_EXP_MAX) return -1.0;
    return r.m * pow2i(r.e);
}

struct B2fState {
    double n0, n1;   // Viterbi nu
    double f0, f1;   // filter
    double b0, b1;   // backward
    double lsum;     // sum log b_1 (log-likelihood)
    int32_t esum, ycur;
};

// Emission of one step: (l0, l1) literal log dbinom values, rho = b_0/b_1 as a double.
template <bool WANT_LB, bool WANT_RHO>
HB_HD void b2f_emit(const Binom2Args &a, const B2fSmem &m, int32_t xv, int32_t nv, double &l0,
                    double &l1, double &rho)
{
    if ((uint32_t)nv <= (uint32_t)m.slut_n && (uint32_t)xv <= (uint32_t)nv) {
        const int32_t ix = ((nv * (nv + 1)) >> 1) + xv;
        if (WANT_LB) {
            const double2 v = m.slb[ix];
            l0 = v.x;
            l1 = v.y;
        }
        if (WANT_RHO) rho = m.srho[ix];
    } else if ((uint32_t)nv <= (uint32_t)a.lut_nmax && (uint32_t)xv <= (uint32_t)nv) {
        const int32_t ix = ((nv * (nv + 1)) >> 1) + xv;
        if (WANT_LB) {
            const double2 v = a.lut_lb[ix];
            l0 = v.x;
            l1 = v.y;
        }
        if (WANT_RHO) rho = a.lut_rho_d[ix];
    } else {
        if (a.c.bb) *a.status = 2;
        RhoME r;
        double sl0, sl1; // (address-taken by the out-of-line call: keep l0/l1 themselves in registers)
        binom2_emit_slow(a.c, xv, nv, sl0, sl1, r);
        l0 = sl0;
        l1 = sl1;
        rho = b2f_rho_double(r);
    }
    if (WANT_RHO && hi32(rho) < 0) { // marked entry: this launch needs the general kernel
        *a.gate = 1;
        rho = 1.0;
    }
}

// phi <- ((phi Pi) . (rho, 1)) 2^-es, es = exponent of the un-normalised sum.  Returns es; rs = rho 2^-es.
HB_HD int32_t b2f_fwd_step(const Binom2Const &c, B2fState &s, double rho, bool first, double &rs,
                           int32_t &ph)
{
    double p0, p1;
    if (first) {
        p0 = c.delta[0];
        p1 = c.delta[1];
    } else {
        p0 = dfma(s.f1, c.Pi[2], s.f0 * c.Pi[0]);
        p1 = dfma(s.f1, c.Pi[3], s.f0 * c.Pi[1]);
    }
    const double q0 = p0 * rho;
    const int32_t es = expo(q0 + p1);
    ph = (1023 - es) << 20;
    const double sc = mk64(ph, 0);
    rs = rho * sc;
    s.f0 = q0 * sc;
    s.f1 = p1 * sc;
    return es;
}

template <bool VIT, bool FB, bool LL>
HB_HD void b2f_forward_one(const Binom2Args &a, const B2fSmem &m, B2fState &s, int32_t xv,
                           int32_t nv, bool first, uint32_t &pk, int j)
{
    double l0 = 0, l1 = 0, rho = 1.0;
    b2f_emit<VIT || LL, FB>(a, m, xv, nv, l0, l1, rho);
    if (VIT) {
        if (first) {
            s.n0 = a.c.logdelta[0] + l0;
            s.n1 = a.c.logdelta[1] + l1;
        } else {
            pk |= binom2_vit_step(a.c, s.n0, s.n1, l0, l1) << (2 * j);
        }
    }
    if (FB) {
        double rs;
        int32_t ph;
        const int32_t es = b2f_fwd_step(a.c, s, rho, first, rs, ph);
        if (LL) {
            s.esum += es;
            s.lsum += l1;
        }
    }
}

HB_HD void b2f_replay_one(const Binom2Args &a, B2fState &s, int32_t xv, int32_t nv, bool first,
                          const B2fSmem &m, int j)
{
    double l0, l1, rho = 1.0, rs;
    int32_t ph;
    b2f_emit<false, true>(a, m, xv, nv, l0, l1, rho);
    b2f_fwd_step(a.c, s, rho, first, rs, ph);
    double2 v;
    v.x = s.f0;
    v.y = rs;
    m.pr[j * m.st] = v;
    m.ph[j * m.st] = ph;
}

template <bool VIT, bool FB>
HB_HD void b2f_backward_one(const Binom2Const &c, B2fState &s, uint8_t *yp, double *g0, double *g1,
                            uint32_t pk, int j, const B2fSmem &m)
{
    if (VIT) {
        st_stream(yp, (uint8_t)(s.ycur + 1));
        s.ycur = (int32_t)((pk >> (2 * j + s.ycur)) & 1u);
    }
    if (FB) {
        const double2 v = m.pr[j * m.st];
        const double sc = mk64(m.ph[j * m.st], 0);
        const double ga = v.x * s.b0;
        st_stream(g0, ga);
        st_stream(g1, fabs(1.0 - ga));
        const double v0 = v.y * s.b0, v1 = sc * s.b1;
        s.b0 = dfma(c.Pi[1], v1, c.Pi[0] * v0);
        s.b1 = dfma(c.Pi[3], v1, c.Pi[2] * v0);
    }
}

// PK: the counts come packed, one word (n << 16) | x per position in a.x (one ring, one 4-byte
// load per step and sweep instead of two).
template <bool PK>
HB_HD void b2f_unpack(int32_t w, int32_t n_sep, int32_t &xv, int32_t &nv)
{
    if (PK) {
        xv = w & 0xffff;
        nv = (int32_t)((uint32_t)w >> 16);
    } else {
        xv = w;
        nv = n_sep;
    }
}

// Warp-cooperative staging on the tiled layout (see xs_refill_half_coop in hb_norm3_fast.cuh): a half
// chunk of a warp's counts is HB_HALF rows of 128 contiguous bytes = 32 pieces of 16 bytes, one
// cp.async per lane.  slots = this chain's slot of the ring's row 0, g = this chain's element of the
// half's first row.
HB_HD void b2f_refill_half_coop(int32_t *slots, int st, int h, const int32_t *g, int lane_c, int lane)
{
#if defined(__CUDA_ARCH__)
    const int r = lane >> 3, e = (lane & 7) * 4;
    unsigned sa = (unsigned)__cvta_generic_to_shared(slots - lane_c + (h * HB_HALF + r) * st + e);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g - lane_c + r * 32 + e) : "memory");
#else
    (void)slots, (void)st, (void)h, (void)g, (void)lane_c, (void)lane;
#endif
}
HB_HD void b2f_warp_sync()
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

// TL: warp-tiled layout known at compile time (strides fold into immediates, as in norm3_fast_chain)
template <bool VIT, bool FB, bool LL, bool PK, bool TL = false>
HB_HD void binom2_fast_chain(const Binom2Args &a, int32_t seg, int64_t b, const B2fSmem &m)
{
    const int32_t tiled = TL ? 1 : a.tiled;
#if defined(__CUDA_ARCH__)
    constexpr bool COOP = TL;
    const int lane = (int)(threadIdx.x & 31u);
#else
    constexpr bool COOP = false;
    const int lane = 0;
#endif
    const int lane_c = (int)(b & 31);
    const Binom2Const &c = a.c;
    const int64_t t0 = a.seg_off[seg];
    const int32_t L = (int32_t)(a.seg_off[seg + 1] - t0);
    if (L <= 0) return;
    const int32_t nfull = L / HB_CHUNK;
    const int32_t tail = L - nfull * HB_CHUNK;
    const int64_t cbase = a.chunk_off[seg];
    const int64_t ld = lay_stride(tiled, a.ld), sB = lay_stride(tiled, a.B);
    const int64_t sy = lay_stride(tiled, a.ldy), sg = lay_stride(tiled, a.ldg);
    const int32_t *xp = a.x + lay_off(tiled, a.T, a.ld, t0, b);
    const int32_t *np = PK ? nullptr : a.n + lay_off(tiled, a.T, a.ld, t0, b);
    // back-pointers: 2 bits / step, one uint16 per chunk
    b2f_psi_t *psi = reinterpret_cast<b2f_psi_t *>(a.psi) + lay_off(tiled, a.nchunks, a.B, cbase, b);
    double *ckpt = a.ckpt + (tiled ? (((b >> 5) * a.nchunks + cbase) * 2 << 5) + (b & 31) : cbase * 2 * a.B + b);
    uint8_t *ybase = a.y ? a.y + lay_off(tiled, a.T, a.ldy, t0, b) : nullptr;
    double *gbase = a.gamma ? a.gamma + lay_off(tiled, a.T, a.ldg, t0, b) : nullptr;

    B2fState s;
    s.n0 = s.n1 = 0;
    s.f0 = s.f1 = 0;
    s.lsum = 0;
    s.esum = 0;

    // ------------------------------------------------------------ forward sweep
    // counts are staged through shared memory with cp.async, one chunk ahead, in half-chunk groups
    // (+ an L2 prefetch HB_B2F_PF_CHUNKS further ahead)
    AsyncRing4 ring;
    if (nfull > 0) {
        const int32_t *xr0 = xp, *nr0 = np;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (COOP) {
                b2f_refill_half_coop(m.xs, m.st, h, xr0, lane_c, lane);
                if (!PK) b2f_refill_half_coop(m.ns, m.st, h, nr0, lane_c, lane);
                xr0 += HB_HALF * ld;
                if (!PK) nr0 += HB_HALF * ld;
            } else {
#pragma unroll
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    cpa4(ring, m.xs + j * m.st, xr0);
                    if (!PK) cpa4(ring, m.ns + j * m.st, nr0);
                    xr0 += ld;
                    if (!PK) nr0 += ld;
                }
            }
            cpa_commit(ring);
        }
    }
    {
        const int32_t *xr = xp + (int64_t)HB_CHUNK * ld, *nr = PK ? nullptr : np + (int64_t)HB_CHUNK * ld;
        b2f_psi_t *pw = psi;
        double *ck = ckpt;
        for (int32_t ch = 0; ch < nfull; ch++) {
            const bool more = ch + 1 < nfull;
            const bool pf = ch + 1 + HB_B2F_PF_CHUNKS < nfull;
            uint32_t pk = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                cpa_wait<1>(ring);
                if (COOP) b2f_warp_sync();
                const int32_t *xh = xr, *nh = nr;
#pragma unroll
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    int32_t xv, nv;
                    b2f_unpack<PK>(m.xs[j * m.st], PK ? 0 : m.ns[j * m.st], xv, nv);
                    if (!COOP && more) {
                        cpa4(ring, m.xs + j * m.st, xr);
                        if (!PK) cpa4(ring, m.ns + j * m.st, nr);
                    }
                    if (TL) {
                        if (j == 0 && pf) {
                            prefetch_l2_chunk(xr + (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * 32, (int)(b & 31), HB_CHUNK);
                            if (!PK)
                                prefetch_l2_chunk(nr + (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * 32, (int)(b & 31),
                                                  HB_CHUNK);
                        }
                    } else if (pf) {
                        prefetch_l2(xr + (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * ld);
                        if (!PK) prefetch_l2(nr + (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * ld);
                    }
                    xr += ld;
                    if (!PK) nr += ld;
                    b2f_forward_one<VIT, FB, LL>(a, m, s, xv, nv, j == 0 && ch == 0, pk, j);
                }
                if (COOP) {
                    b2f_warp_sync(); // every lane has read this half's rows
                    if (more) {
                        b2f_refill_half_coop(m.xs, m.st, h, xh, lane_c, lane);
                        if (!PK) b2f_refill_half_coop(m.ns, m.st, h, nh, lane_c, lane);
                    }
                }
                cpa_commit(ring);
            }
            if (VIT) *pw = (b2f_psi_t)pk;
            pw += sB;
            ck += 2 * sB;
            if (FB && (more || tail > 0)) {
                ck[0] = s.f0;
                ck[sB] = s.f1;
            }
        }
    }
    cpa_wait<0>(ring);
    if (tail > 0) {
        uint32_t pk = 0;
        for (int j = 0; j < tail; j++) {
            const int64_t o = (int64_t)(nfull * HB_CHUNK + j) * ld;
            int32_t xv, nv;
            b2f_unpack<PK>(xp[o], PK ? 0 : np[o], xv, nv);
            b2f_forward_one<VIT, FB, LL>(a, m, s, xv, nv, nfull == 0 && j == 0, pk, j);
        }
        if (VIT) psi[(int64_t)nfull * sB] = (b2f_psi_t)pk;
    }

    if (FB && LL) {
        // LL = log(phi_0 + phi_1) + ln2 * esum + sum_t log b_1
        const double LN2 = 0.693147180559945309417232121458;
        a.ll[(int64_t)seg * a.B + b] = (log(s.f0 + s.f1) + LN2 * (double)s.esum) + s.lsum;
        if (!VIT && !a.gamma) return; // HB_WANT_LL alone: the forward sweep is the whole job
    }
    s.ycur = 0;
    if (VIT) {
        if (s.n0 == -INFINITY || s.n1 == -INFINITY) *a.status = 1;
        if (s.n1 > s.n0) s.ycur = 1;
    }

    // ------------------------------------------------------------ backward sweep
    if (FB) {
        const double q = 1.0 / (s.f0 + s.f1);
        s.b0 = s.b1 = q;
    }
    if (FB && nfull > 0) {
        const int32_t *xl = xp + (int64_t)(nfull - 1) * HB_CHUNK * ld;
        const int32_t *nl = PK ? nullptr : np + (int64_t)(nfull - 1) * HB_CHUNK * ld;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (COOP) {
                b2f_refill_half_coop(m.xs, m.st, h, xl, lane_c, lane);
                if (!PK) b2f_refill_half_coop(m.ns, m.st, h, nl, lane_c, lane);
                xl += HB_HALF * ld;
                if (!PK) nl += HB_HALF * ld;
            } else {
#pragma unroll
                for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                    cpa4(ring, m.xs + j * m.st, xl);
                    if (!PK) cpa4(ring, m.ns + j * m.st, nl);
                    xl += ld;
                    if (!PK) nl += ld;
                }
            }
            cpa_commit(ring);
        }
    }
    uint32_t pk_nx = 0;
    double c0_nx = 0, c1_nx = 0;
    if (nfull > 0) {
        if (VIT) pk_nx = psi[(int64_t)(nfull - 1) * sB];
        if (FB && nfull > 1) {
            const double *ck = ckpt + (int64_t)(nfull - 1) * 2 * sB;
            c0_nx = ck[0];
            c1_nx = ck[sB];
        }
    }
    const int64_t plane = a.gplane;
    if (tail > 0) {
        uint32_t pk = 0;
        if (VIT) pk = psi[(int64_t)nfull * sB];
        if (FB) {
            if (nfull > 0) {
                const double *ck = ckpt + (int64_t)nfull * 2 * sB;
                s.f0 = ck[0];
                s.f1 = ck[sB];
            }
            for (int j = 0; j < tail; j++) {
                const int64_t o = (int64_t)(nfull * HB_CHUNK + j) * ld;
                int32_t xv, nv;
                b2f_unpack<PK>(xp[o], PK ? 0 : np[o], xv, nv);
                b2f_replay_one(a, s, xv, nv, nfull == 0 && j == 0, m, j);
            }
        }
        for (int j = tail - 1; j >= 0; j--) {
            const int64_t rel = nfull * HB_CHUNK + j;
            double *gr = FB ? gbase + rel * sg : nullptr;
            b2f_backward_one<VIT, FB>(c, s, VIT ? ybase + rel * sy : nullptr, gr, gr + plane, pk,
                                      j, m);
        }
    }
    {
        const int64_t rell = (int64_t)nfull * HB_CHUNK - 1;
        uint8_t *yp = VIT ? ybase + rell * sy : nullptr;
        double *g0 = FB ? gbase + rell * sg : nullptr;
        double *g1 = g0 + plane;
        // refill pointers: the chunk before the one being replayed
        const int32_t *xr = xp + (int64_t)(nfull - 2) * HB_CHUNK * ld;
        const int32_t *nr = PK ? nullptr : np + (int64_t)(nfull - 2) * HB_CHUNK * ld;
        const b2f_psi_t *pr = psi + (int64_t)(nfull - 2) * sB;
        const double *cr = ckpt + (int64_t)(nfull - 2) * 2 * sB;
        for (int32_t ch = nfull - 1; ch >= 0; ch--) {
            const uint32_t pk = pk_nx;
            if (FB && ch > 0) {
                s.f0 = c0_nx;
                s.f1 = c1_nx;
            }
            if (ch > 0) {
                if (VIT) pk_nx = *pr;
                if (FB && ch > 1) {
                    c0_nx = cr[0];
                    c1_nx = cr[sB];
                }
            }
            if (ch > 1 + HB_B2F_PF_CHUNKS) { // checkpoints / back-pointers further back
                if (VIT) prefetch_l2(pr - (int64_t)HB_B2F_PF_CHUNKS * sB);
                if (FB) {
                    prefetch_l2(cr - (int64_t)HB_B2F_PF_CHUNKS * 2 * sB);
                    prefetch_l2(cr - (int64_t)HB_B2F_PF_CHUNKS * 2 * sB + sB);
                }
            }
            pr -= sB;
            cr -= 2 * sB;
            if (FB) {
                const bool more = ch > 0;
                const bool pf = ch > 1 + HB_B2F_PF_CHUNKS;
                const int32_t *xq = xr, *nq = nr;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    cpa_wait<1>(ring);
                    if (COOP) b2f_warp_sync();
                    const int32_t *xh = xq, *nh = nq;
#pragma unroll
                    for (int j = h * HB_HALF; j < (h + 1) * HB_HALF; j++) {
                        int32_t xv, nv;
                        b2f_unpack<PK>(m.xs[j * m.st], PK ? 0 : m.ns[j * m.st], xv, nv);
                        if (!COOP && more) {
                            cpa4(ring, m.xs + j * m.st, xq);
                            if (!PK) cpa4(ring, m.ns + j * m.st, nq);
                        }
                        if (TL) {
                            if (j == 0 && pf) {
                                prefetch_l2_chunk(xq - (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * 32, (int)(b & 31),
                                                  HB_CHUNK);
                                if (!PK)
                                    prefetch_l2_chunk(nq - (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * 32, (int)(b & 31),
                                                      HB_CHUNK);
                            }
                        } else if (pf) {
                            prefetch_l2(xq - (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * ld);
                            if (!PK) prefetch_l2(nq - (int64_t)HB_B2F_PF_CHUNKS * HB_CHUNK * ld);
                        }
                        xq += ld;
                        if (!PK) nq += ld;
                        b2f_replay_one(a, s, xv, nv, j == 0 && ch == 0, m, j);
                    }
                    if (COOP) {
                        b2f_warp_sync();
                        if (more) {
                            b2f_refill_half_coop(m.xs, m.st, h, xh, lane_c, lane);
                            if (!PK) b2f_refill_half_coop(m.ns, m.st, h, nh, lane_c, lane);
                        }
                    }
                    cpa_commit(ring);
                }
                xr -= (int64_t)HB_CHUNK * ld;
                if (!PK) nr -= (int64_t)HB_CHUNK * ld;
            }
#pragma unroll
            for (int j = HB_CHUNK - 1; j >= 0; j--) {
                b2f_backward_one<VIT, FB>(c, s, yp, g0, g1, pk, j, m);
                if (VIT) yp -= sy;
                if (FB) {
                    g0 -= sg;
                    g1 -= sg;
                }
            }
        }
    }
    cpa_wait<0>(ring);
}

} // namespace hb
