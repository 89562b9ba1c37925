// hb_device.cuh — small device helpers shared by the HMM kernels (sm_100a).
//
// Everything here is `HB_HD` so that tests/emul can compile the SAME chain bodies for the host and
// step through them without a GPU (a debugging aid for the CPU test-suite; the package never
// loads that build and has no CPU path).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define HB_HD inline
#define HB_HD_NOINLINE inline
struct double2 {
    double x, y;
};
#endif

namespace hb {

// ---------------------------------------------------------------- bit access
HB_HD int32_t hi32(double v)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return (int32_t)(u >> 32);
#endif
}
HB_HD int32_t lo32(double v)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return (int32_t)(u & 0xffffffffu);
#endif
}
HB_HD double mk64(int32_t hi, int32_t lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}
// Unbiased binary exponent of a positive normal double.
HB_HD int32_t expo(double v) { return ((hi32(v) >> 20) & 0x7ff) - 1023; }
// v * 2^k by integer add on the exponent field (v normal, result must stay normal).
HB_HD double scale2(double v, int32_t k) { return mk64(hi32(v) + (k << 20), lo32(v)); }
// 2^k as a double, k in [-1022, 1023].
HB_HD double pow2i(int32_t k) { return mk64((1023 + k) << 20, 0); }

HB_HD double dfma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// ---------------------------------------------------------------- layouts
// element offset of (position t, column b) and the stride between consecutive positions
HB_HD int64_t lay_off(int32_t tiled, int64_t T, int64_t ld, int64_t t, int64_t b)
{
    return tiled ? (((b >> 5) * T + t) << 5) + (b & 31) : t * ld + b;
}
HB_HD int64_t lay_stride(int32_t tiled, int64_t ld) { return tiled ? 32 : ld; }

// ---------------------------------------------------------------- streaming loads / stores
// The chain kernels stream their inputs and outputs exactly once per sweep: keep them out of L1
// (evict-first) so the emission table and the constant data stay resident there.
template <typename T>
HB_HD T ld_stream(const T *p)
{
#if defined(__CUDA_ARCH__)
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T>
HB_HD void st_stream(T *p, T v)
{
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}

// Experiment (off by default, -DHB_SCRATCH_L2_HINTS=1): the scratch written in the forward sweep and read back
// once in the backward sweep (back-pointer words, filter checkpoints) with an L2 evict-LAST policy on the stores
// and evict-first on the one re-read, so that the ~100 MB of live scratch might survive in the 126 MB L2 between
// the sweeps.  Measured A/B in one session (round 2, tools/build_variant.sh): SLOWER — Viterbi + posteriors
// 1.800 vs 1.771 ms (median), posteriors only 1.589 vs 1.521 ms: pinning the scratch takes L2 away from the
// read-ahead of x and from write combining of the posterior stores.
#ifndef HB_SCRATCH_L2_HINTS
#define HB_SCRATCH_L2_HINTS 0
#endif
struct L2Pol {
    uint64_t keep, drop;
};
HB_HD L2Pol l2pol_make()
{
    L2Pol p;
    p.keep = p.drop = 0;
#if defined(__CUDA_ARCH__) && HB_SCRATCH_L2_HINTS
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.drop));
#endif
    return p;
}
template <typename T>
HB_HD void st_keep(const L2Pol &p, T *ptr, T v)
{
#if defined(__CUDA_ARCH__) && HB_SCRATCH_L2_HINTS
    static_assert(sizeof(T) == 8, "8-byte scratch words");
    uint64_t u;
    memcpy(&u, &v, 8);
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(ptr), "l"(u), "l"(p.keep) : "memory");
#else
    (void)p;
    *ptr = v;
#endif
}
template <typename T>
HB_HD T ld_drop(const L2Pol &p, const T *ptr)
{
#if defined(__CUDA_ARCH__) && HB_SCRATCH_L2_HINTS
    static_assert(sizeof(T) == 8, "8-byte scratch words");
    uint64_t u;
    asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(u) : "l"(ptr), "l"(p.drop) : "memory");
    T v;
    memcpy(&v, &u, 8);
    return v;
#else
    (void)p;
    return *ptr;
#endif
}

// Warp-tiled layout: one instruction prefetches a whole chunk.  `p` is this lane's element of the
// chunk's first row; the chunk is `lines` consecutive 128-byte lines starting at the warp's row, and
// lane l touches line l % lines (instead of every lane touching its own row's two lines at every step).
template <typename T>
HB_HD void prefetch_l2_chunk(const T *p, int lane, int lines)
{
#if defined(__CUDA_ARCH__)
    const char *base = reinterpret_cast<const char *>(p - lane) + 128 * (lane % lines);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base));
#else
    (void)p, (void)lane, (void)lines;
#endif
}

// Hint: bring the line holding *p into L2 (no register, no dependency).
HB_HD void prefetch_l2(const void *p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Hint: bring the line holding *p into L1 (few-thread sequential walkers: nothing else hides L2 latency).
HB_HD void prefetch_l1(const void *p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---------------------------------------------------------------- cp.async staging of 4-byte elements
// Per-thread asynchronous copies global -> shared with explicit commit / wait groups.  Unlike a
// register prefetch (whose first use can end up waiting on the scoreboard it shares with the NEXT
// prefetch), a wait_group only waits for the groups it names.  The host build defers visibility
// to the wait, so the CPU emulation catches group-accounting mistakes.
struct AsyncRing4 {
#if !defined(__CUDA_ARCH__)
    int32_t pend_v[64];
    int32_t *pend_p[64];
    int pend_g[64];
    int npend = 0, cur_group = 0;
#endif
};
HB_HD void cpa4(AsyncRing4 &r, int32_t *dst_smem, const int32_t *g)
{
#if defined(__CUDA_ARCH__)
    unsigned sa = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
#else
    r.pend_v[r.npend] = *g;
    r.pend_p[r.npend] = dst_smem;
    r.pend_g[r.npend] = r.cur_group;
    r.npend++;
#endif
}
HB_HD void cpa_commit(AsyncRing4 &r)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    r.cur_group++;
#endif
}
// Wait until at most N of the most recently committed groups are still in flight.
template <int N>
HB_HD void cpa_wait(AsyncRing4 &r)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#else
    int keep = 0;
    for (int i = 0; i < r.npend; i++) {
        if (r.pend_g[i] < r.cur_group - N) {
            *r.pend_p[i] = r.pend_v[i];
        } else {
            r.pend_v[keep] = r.pend_v[i];
            r.pend_p[keep] = r.pend_p[i];
            r.pend_g[keep] = r.pend_g[i];
            keep++;
        }
    }
    r.npend = keep;
#endif
}

// ---------------------------------------------------------------- bulk (TMA) staging on the tiled layout
// On the warp-tiled layout the rows a warp needs for one chunk are ONE contiguous 2 KB block, so one
// lane moves them with a single cp.async.bulk (the TMA engine's non-tensor form) instead of 32 lanes
// x 8 cp.async, completion tracked by an mbarrier per stage.  Two stages per warp: while chunk c is
// consumed, chunk c+1 is in flight; chunk c+2 is requested as soon as the warp is done with c.
struct BulkRing {
    double *buf;   // this lane's view of the warp's staging area: element (stage s, row j) at buf[(s*CH + j)*32]
    uint32_t par[2];
#if defined(__CUDACC__)
    uint32_t mbar[2], sbuf; // shared-space addresses: mbarriers, staging area of the warp (lane 0)
    uint32_t wmask;
    bool leader;
#else
    double hbuf[64];        // host emulation (g++ build): one lane, copies are immediate
#endif
};

template <int CH>
HB_HD void bulk_issue(BulkRing &r, int s, const double *lane_src /* this lane's element of the chunk's first row */)
{
#if defined(__CUDA_ARCH__)
    if (r.leader) {
        const uint32_t bytes = CH * 32 * 8;
        const double *src = lane_src - (threadIdx.x & 31);
        const uint32_t dst = r.sbuf + (uint32_t)s * bytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(r.mbar[s]), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(r.mbar[s]) : "memory");
    }
#elif !defined(__CUDACC__)
    for (int j = 0; j < CH; j++) r.hbuf[s * CH + j] = lane_src[(int64_t)j * 32];
#endif
}
HB_HD void bulk_wait(BulkRing &r, int s)
{
#if defined(__CUDA_ARCH__)
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(r.mbar[s]), "r"(r.par[s]) : "memory");
    }
#endif
    r.par[s] ^= 1u;
}
template <int CH>
HB_HD double bulk_read(const BulkRing &r, int s, int j)
{
#if defined(__CUDACC__)
    return r.buf[(s * CH + j) * 32];
#else
    return r.hbuf[s * CH + j];
#endif
}
// every lane is done reading stage s (before the leader overwrites it)
HB_HD void bulk_release(BulkRing &r)
{
#if defined(__CUDA_ARCH__)
    __syncwarp(r.wmask);
#endif
}

// ---------------------------------------------------------------- exact division by a chain constant
// z = RN(d / b) given y = RN(1/b): Markstein's correction (q = d*y; r = d - b*q exactly by FMA;
// q' = q + r*y) is correctly rounded for every d when y is the correctly rounded reciprocal
// (tests/test_numerics.py checks the hardest cases).  3 fp64 instructions instead of ~12.
HB_HD double div_by_const(double d, double b, double y)
{
    double q = d * y;
    double r = dfma(-b, q, d);
    return dfma(r, y, q);
}

// ---------------------------------------------------------------- reciprocal to ~1e-16 (posterior normalisation)
HB_HD double rcp_fast(double a)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#else
    double r = (double)(1.0f / (float)a);
#endif
    double e = dfma(-a, r, 1.0);
    r = dfma(r, e, r);
    e = dfma(-a, r, 1.0);
    r = dfma(r, e, r);
    return r;
}

// ---------------------------------------------------------------- exp(u) and exp(-u) together
// Returns Pp = e^r, Pm = e^-r with u = n ln2 + r, |r| <= ln2/2, and n clamped to +-NCLAMP so that
// exponent arithmetic on the results can never leave the normal range (single-position evidence
// saturates at e^+-346; Gaussian ratios on this path stay far below).  Relative error < 3e-16.
// Shared range reduction, even/odd split of the degree-13 Taylor polynomial: 4 + 15 + 2 fp64 ops.
#define HB_NCLAMP 500
HB_HD void exp_pair(double u, double &Pp, double &Pm, int32_t &n)
{
    const double L2E = 1.4426950408889634074;
    const double MAGIC = 6755399441055744.0; // 1.5 * 2^52
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double t = dfma(u, L2E, MAGIC);
    int32_t ni = lo32(t);
    double nf = t - MAGIC;
    double r = dfma(nf, -LN2_HI, u);
    r = dfma(nf, -LN2_LO, r);
    double q = r * r;
    double ev = dfma(q, 1.0 / 479001600.0, 1.0 / 3628800.0); // q^6/12! ... 1
    ev = dfma(q, ev, 1.0 / 40320.0);
    ev = dfma(q, ev, 1.0 / 720.0);
    ev = dfma(q, ev, 1.0 / 24.0);
    ev = dfma(q, ev, 0.5);
    ev = dfma(q, ev, 1.0);
    double od = dfma(q, 1.0 / 6227020800.0, 1.0 / 39916800.0); // r*(q^6/13! ... 1)
    od = dfma(q, od, 1.0 / 362880.0);
    od = dfma(q, od, 1.0 / 5040.0);
    od = dfma(q, od, 1.0 / 120.0);
    od = dfma(q, od, 1.0 / 6.0);
    od = dfma(q, od, 1.0);
    od = od * r;
    Pp = ev + od;
    Pm = ev - od;
    n = ni > HB_NCLAMP ? HB_NCLAMP : (ni < -HB_NCLAMP ? -HB_NCLAMP : ni);
}

// exp(u) alone (general, non-symmetric means): Pp * 2^n.
HB_HD void exp_one(double u, double &Pp, int32_t &n)
{
    double Pm;
    exp_pair(u, Pp, Pm, n);
}

} // namespace hb
