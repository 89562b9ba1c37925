// hb_norm3_resident.cuh — 3-state Gaussian chain, posteriors only, CHAIN-RESIDENT and parallel along the
// gene axis: the design north_star sketches ("a warp-level associative scan along the gene axis") and the
// round-1 review asked to be built and measured against the thread-per-chain kernel.
//
// Semantics: HiddenMarkov::forwardback + exp(logalpha + logbeta - LL) for the reference's
// symmetric parameter block (R/HoneyBADGER.R:1144-1150), tolerance 1e-9 — the same posteriors
// hb_norm3_fast.cuh produces, by a different route.  Viterbi is NOT here: its recursion is sequential per
// chain (rounded max-plus, first index wins), and a resident design cannot keep enough chains on chip to
// hide that latency — see DESIGN.md 4.2 for the measurement.
//
// One chain = one block of NW warps (P = 32 NW lanes); the chain's L positions are cut into P contiguous
// pieces of m positions, one per lane.  x is read ONCE, posteriors are written ONCE: 32 B per step of HBM
// traffic instead of the 46 of the two-sweep kernel (no second read of x, no checkpoints).
//
//   phase 0  all lanes, positions strided across lanes (coalesced 256 B loads): u = A x + B,
//            E0 = K e^-u, E2 = K e^u (b_k / b_1, one exp pair per position) -> shared memory
//   phase 1  lane: transfer matrix of its piece, M = prod_t (Pi diag(E0_t, 1, E2_t)), 3 row recursions of
//            8 fp64 each (Pi = (a - t) I + t J: a row times Pi is 3 FMA + the row sum)
//   phase 1b block: inclusive prefix and suffix products of the M across lanes (Kogge-Stone over shuffles,
//            warps stitched through shared memory); all terms are non-negative, so the products are
//            associative up to ~P S u relative error, no cancellation
//   phase 2  lane: filter forward through its piece from the entry vector alpha0 M_0 .. M_{b-1}, rescaled by
//            2^-es_t per step as in hb_norm3_fast.cuh; parks phi_0, phi_2, the scaled ratios and 2^-es_t
//   phase 3  lane: beta backwards through its piece from the exit vector M_{b+1} .. M_{P-1} 1, normalised so
//            that sum_k phi(k) beta(k) = 1 at the piece's last position; using the SAME 2^-es_t keeps that sum
//            at 1, hence gamma_t(k) = phi_t(k) beta_t(k) (gamma_1 = |1 - gamma_0 - gamma_2|) — written over
//            the parked phi in place
//   phase 4  all lanes, strided: the three posterior planes leave as coalesced 256 B stores
//
// Layout: CHAIN-MAJOR, x[b * T + t] — R's own column-major [T x B] — and gamma[(k * B + b) * T + t] = R's
// [T x B x 3]; a chain's positions must be contiguous for phase 0 / 4 to coalesce (the warp-tiled layout of
// the thread-per-chain kernels puts consecutive positions of a chain 256 B apart).
//
// Shared memory: 36 B per position (E0, E2, phi_0 / gamma_0, phi_2 / gamma_2, 2^-es) — what bounds the
// number of resident chains per SM, i.e. the occupancy.
#pragma once
#include "hb_norm3_fast.cuh"

namespace hb {

#define HB_FBR_MAX_NW 4 // warps per chain: 1, 2 or 4

struct FbrArgs {
    const double *x;      // chain-major x[b * T + t]
    int64_t T, B;
    int64_t t0;           // segment start
    int32_t L;            // segment length
    int32_t m;            // positions per lane (odd: conflict-free 8-byte shared-memory accesses at stride m)
    int32_t seg;
    int32_t sd_per_seg;
    const double *sd;     // [B] or [nseg * B]
    double *gamma;        // gamma[(k * B + b) * T + t]
    int32_t *gate;        // set to 1 when a piece's transfer matrix left the double range (caller re-runs
                          // the launch on the thread-per-chain kernel)
    Norm3Const c;
};

#if defined(__CUDACC__)

struct Mat3 {
    double a[9]; // row-major
};

// C = A B (non-negative entries), then scaled by a power of two so that the entry sum is in [1, 2)
__device__ __forceinline__ Mat3 fbr_mul(const Mat3 &A, const Mat3 &Bm)
{
    Mat3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++)
            C.a[3 * i + k] = dfma(A.a[3 * i], Bm.a[k], dfma(A.a[3 * i + 1], Bm.a[3 + k], A.a[3 * i + 2] * Bm.a[6 + k]));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) s += C.a[i];
    const double sc = pow2i(-expo(s)); // s > 0 normal unless the matrices degenerated (gated in phase 1)
#pragma unroll
    for (int i = 0; i < 9; i++) C.a[i] *= sc;
    return C;
}

__device__ __forceinline__ Mat3 fbr_shfl(const Mat3 &A, int src, bool up)
{
    Mat3 R;
#pragma unroll
    for (int i = 0; i < 9; i++)
        R.a[i] = up ? __shfl_up_sync(0xffffffffu, A.a[i], src) : __shfl_down_sync(0xffffffffu, A.a[i], src);
    return R;
}

// One chain per block.  smem: E0[Lp] E2[Lp] F0[Lp] F2[Lp] (double) PH[Lp] (int32), Lp = P * m; the scan's
// warp totals (2 * NW matrices) live at the front of F0 until phase 2 overwrites them.
template <int NW>
__device__ __forceinline__ void fbr_chain(const FbrArgs &a, int64_t b, double *smem)
{
    constexpr int P = 32 * NW;
    const Norm3Const &c = a.c;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, m = a.m, Lp = P * m;
    double *E0 = smem, *E2 = E0 + Lp, *F0 = E2 + Lp, *F2 = F0 + Lp;
    int32_t *PH = reinterpret_cast<int32_t *>(F2 + Lp);
    const int64_t pix = a.sd_per_seg ? (int64_t)a.seg * a.B + b : b;
    const double isd = 1.0 / a.sd[pix]; // IEEE division: the value the host-side preparation of the other kernels uses
    const double mm = c.mean[2] - c.mean[1], mi = mm * isd;
    const double A = mi * isd, Bc = -A * c.mean[1];
    const double K = exp(-0.5 * mi * mi);
    const double *xp = a.x + b * a.T + a.t0;

    // ---- phase 0: emission ratios of every position
    for (int t = tid; t < L; t += P) {
        const double xv = ld_stream(xp + t);
        double Pp, Pm;
        int32_t n;
        fast_exp_pair(c, dfma(xv, A, Bc), Pp, Pm, n);
        E0[t] = scale2(Pm * K, -n);
        E2[t] = scale2(Pp * K, n);
    }
    __syncthreads();

    // ---- phase 1: this lane's piece [s, e) -> transfer matrix
    const int s = min(tid * m, L), e = min(s + m, L);
    Mat3 M;
#pragma unroll
    for (int i = 0; i < 9; i++) M.a[i] = (i % 4 == 0) ? 1.0 : 0.0;
    const double amt = c.amt, tt = c.tt;
    for (int t = s; t < e; t++) {
        const double e0 = E0[t], e2 = E2[t];
        if (t == 0) { // the chain's first position: phi_0 = delta . b_0, no transition
            M.a[0] = e0;
            M.a[8] = e2;
            continue;
        }
        double tot = 0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double v0 = M.a[3 * r], v1 = M.a[3 * r + 1], v2 = M.a[3 * r + 2];
            const double sm = (v0 + v1) + v2;
            const double ts = tt * sm;
            M.a[3 * r + 1] = dfma(amt, v1, ts);
            M.a[3 * r] = dfma(amt, v0, ts) * e0;
            M.a[3 * r + 2] = dfma(amt, v2, ts) * e2;
            tot += sm;
        }
        if (((t - s) & 7) == 7) { // keep the entries in range: divide by the power of two of the last total
            const double sc = pow2i(-expo(tot));
#pragma unroll
            for (int i = 0; i < 9; i++) M.a[i] *= sc;
        }
    }
    {
        double tot = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) tot += M.a[i];
        if (!(tot > 0.0) || !(tot < 1.7e308)) *a.gate = 1; // degenerate: the caller re-runs this launch
        const double sc = (tot > 0.0 && tot < 1.7e308) ? pow2i(-expo(tot)) : 1.0;
#pragma unroll
        for (int i = 0; i < 9; i++) M.a[i] *= sc;
    }

    // ---- phase 1b: inclusive prefix products Q_b = M_0 .. M_b and suffix products R_b = M_b .. M_{P-1}
    Mat3 Q = M, R = M;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Mat3 qo = fbr_shfl(Q, d, true), ro = fbr_shfl(R, d, false);
        if (lane >= d) Q = fbr_mul(qo, Q);
        if (lane + d < 32) R = fbr_mul(R, ro);
    }
    double av[3], bv[3]; // entry vector of the filter, exit vector of beta (both up to scale)
    {
        Mat3 Qx = fbr_shfl(Q, 1, true), Rx = fbr_shfl(R, 1, false); // exclusive forms inside the warp
        if (lane == 0)
#pragma unroll
            for (int i = 0; i < 9; i++) Qx.a[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (lane == 31)
#pragma unroll
            for (int i = 0; i < 9; i++) Rx.a[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (NW > 1) {
            // warp totals through shared memory (front of F0: not in use before phase 2)
            Mat3 *tq = reinterpret_cast<Mat3 *>(F0), *tr = tq + NW;
            if (lane == 31) tq[warp] = Q;
            if (lane == 0) tr[warp] = R;
            __syncthreads();
            for (int w = warp - 1; w >= 0; w--) Qx = fbr_mul(tq[w], Qx);
            for (int w = warp + 1; w < NW; w++) Rx = fbr_mul(Rx, tr[w]);
            __syncthreads();
        }
        // alpha entry = delta Qx (the first position's matrix already carries "no transition"), beta exit = Rx 1
#pragma unroll
        for (int k = 0; k < 3; k++) {
            av[k] = dfma(c.delta[0], Qx.a[k], dfma(c.delta[1], Qx.a[3 + k], c.delta[2] * Qx.a[6 + k]));
            bv[k] = (Rx.a[3 * k] + Rx.a[3 * k + 1]) + Rx.a[3 * k + 2];
        }
    }

    // ---- phase 2: filter forward through the piece
    double f0 = av[0], f1 = av[1], f2 = av[2];
    for (int t = s; t < e; t++) {
        const double e0 = E0[t], e2 = E2[t];
        double p0, p1, p2;
        int32_t es;
        if (t == 0) {
            p0 = c.delta[0];
            p1 = c.delta[1];
            p2 = c.delta[2];
            es = 0;
        } else {
            const double sm = (f0 + f1) + f2;
            es = expo(sm);
            const double ts = tt * sm;
            p1 = dfma(amt, f1, ts);
            p0 = dfma(amt, f0, ts);
            p2 = dfma(amt, f2, ts);
        }
        const int32_t ph = (1023 - es) << 20;
        const double e0s = scale2(e0, -es), e2s = scale2(e2, -es);
        f1 = p1 * mk64(ph, 0);
        f0 = p0 * e0s;
        f2 = p2 * e2s;
        E0[t] = e0s;
        E2[t] = e2s;
        F0[t] = f0;
        F2[t] = f2;
        PH[t] = ph;
    }

    // ---- phase 3: beta backwards, posteriors in place
    if (e > s) {
        const double q = 1.0 / dfma(f0, bv[0], dfma(f1, bv[1], f2 * bv[2]));
        double b0 = bv[0] * q, b1 = bv[1] * q, b2 = bv[2] * q;
        for (int t = e - 1; t >= s; t--) {
            const double ga = F0[t] * b0, gc = F2[t] * b2;
            F0[t] = ga;
            F2[t] = gc;
            const double v0 = E0[t] * b0, v2 = E2[t] * b2, v1 = mk64(PH[t], 0) * b1;
            const double tsw = tt * ((v0 + v2) + v1);
            b1 = dfma(amt, v1, tsw);
            b0 = dfma(amt, v0, tsw);
            b2 = dfma(amt, v2, tsw);
        }
    }
    __syncthreads();

    // ---- phase 4: coalesced stores of the three planes
    double *g0 = a.gamma + b * a.T + a.t0, *g1 = g0 + a.B * a.T, *g2 = g1 + a.B * a.T;
    for (int t = tid; t < L; t += P) {
        const double ga = F0[t], gc = F2[t];
        st_stream(g0 + t, ga);
        st_stream(g2 + t, gc);
        st_stream(g1 + t, fabs(1.0 - (ga + gc)));
    }
}

#endif // __CUDACC__

// bytes of shared memory per block for a segment of L positions on NW warps; *m = positions per lane (odd)
inline size_t fbr_smem_bytes(int L, int NW, int *m_out)
{
    const int P = 32 * NW;
    int m = (L + P - 1) / P;
    if (m < 3) m = 3;
    if ((m & 1) == 0) m++;
    *m_out = m;
    return (size_t)P * m * 36;
}

} // namespace hb
