// hb_kernels.cu — __global__ entry points and launchers (sm_100a).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
// (-fmad=false: the Viterbi path must keep one rounding per reference operation).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "hb_binom2.cuh"
#include "hb_binom2_fast.cuh"
#include "hb_group.cuh"
#include "hb_inputs.cuh"
#include "hb_kernels.h"
#include "hb_longchain.cuh"
#include "hb_norm3.cuh"
#include "hb_norm3_fast.cuh"
#include "hb_norm3_resident.cuh"
#include "hb_x87.cuh"
#include "hb_x87_pair.cuh"

namespace hb {

// =====================================================================================
// HMM chains: one thread per (cell column, segment); block = HB_BLOCK consecutive cells of
// one segment, segments issued longest first.
// =====================================================================================
#ifndef HB_NORM3_MINB
#define HB_NORM3_MINB 4
#endif
template <bool SYM, bool VIT, bool FB, bool LL>
__global__ void __launch_bounds__(HB_BLOCK, HB_NORM3_MINB)
    norm3_kernel(const __grid_constant__ Norm3Args a)
{
    extern __shared__ double2 smem2[];
    const int32_t seg = a.seg_order[blockIdx.x / a.nblk_per_seg];
    const int64_t b = (int64_t)(blockIdx.x % a.nblk_per_seg) * HB_BLOCK + threadIdx.x;
    if (b >= a.B) return;
    Norm3Smem m;
    m.st = HB_BLOCK;
    m.va = smem2 + threadIdx.x;
    m.vb = m.va + HB_CHUNK * HB_BLOCK;
    m.e2 = reinterpret_cast<double *>(smem2 + 2 * HB_CHUNK * HB_BLOCK) + threadIdx.x;
    m.xs = FB ? m.e2 + HB_CHUNK * HB_BLOCK : reinterpret_cast<double *>(smem2) + threadIdx.x;
    norm3_chain<SYM, VIT, FB, LL>(a, seg, b, m);
}

// Speculate-and-verify path for the reference's symmetric parameter block (hb_norm3_fast.cuh).
// 3 resident blocks/SM (<= 168 registers): measured faster than 4 blocks at <= 128 registers, the
// extra registers buy instruction-level parallelism (tools/build_variant.sh sweeps).
// Also measured: the two sweeps as two kernels (forward + verification, backward; 124-128 registers each
// without spills, 4 blocks/SM, per-chain hand-off of 1 / sum phi_T and the final state): 2.24 ms against
// 1.95 ms fused — blocks in different sweeps overlap read-heavy and write-heavy phases, two uniform
// kernels do not.
#ifndef HB_FAST_MINB
#define HB_FAST_MINB 3
#endif
template <bool VIT, bool FB, bool LL, bool BULK, bool TL>
__global__ void __launch_bounds__(HB_BLOCK, HB_FAST_MINB)
    norm3_fast_kernel(const __grid_constant__ Norm3Args a)
{
    extern __shared__ double2 smem2[];
    const int32_t seg = a.seg_order[blockIdx.x / a.nblk_per_seg];
    int64_t b = (int64_t)(blockIdx.x % a.nblk_per_seg) * HB_BLOCK + threadIdx.x;
    int tix = (int)threadIdx.x;
    bool live = true;
    if (TL && !BULK) {
        // cooperative staging needs complete warps: in a partially filled tile the lanes past the last
        // chain repeat that chain (same slots, same addresses, same values)
        if ((b & ~(int64_t)31) >= a.B) return;
        if (b >= a.B) {
            tix -= (int)(b - (a.B - 1));
            b = a.B - 1;
            live = false;
        }
    } else if (b >= a.B) {
        return;
    }
    Norm3Smem m;
    m.st = HB_BLOCK;
    m.va = smem2 + tix;
    m.vb = m.va + HB_CHUNK * HB_BLOCK;
    // (Viterbi only: no parking area — the block needs just the x ring and the state tiles, 9 KB instead of
    // 45, which leaves the SM's shared memory to whatever runs beside it)
    char *rest = reinterpret_cast<char *>(smem2 + (FB ? 2 * HB_CHUNK * HB_BLOCK : 0));
    m.e2 = reinterpret_cast<double *>(rest) + tix;
    if (FB && !HB_FAST_NOPHI1) rest += sizeof(double) * HB_CHUNK * HB_BLOCK;
    m.ph = reinterpret_cast<int32_t *>(rest) + tix;
    if (FB) rest += sizeof(int32_t) * HB_CHUNK * HB_BLOCK;
    m.xs = reinterpret_cast<double *>(rest) + tix; // cp.async ring (one chunk) ...
    m.ys = reinterpret_cast<uint8_t *>(rest + sizeof(double) * HB_CHUNK * HB_BLOCK) + (tix >> 5) * (HB_CHUNK * 32) +
           (tix & 31); // ... then the warps' [HB_CHUNK][32] state tiles
    BulkRing br;
    br.par[0] = br.par[1] = 0u;
    if (BULK) { // ... or, per warp, two chunks of bulk staging followed by the block's mbarriers
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double *stage = reinterpret_cast<double *>(rest) + (size_t)warp * 2 * HB_CHUNK * 32;
        uint64_t *bars = reinterpret_cast<uint64_t *>(rest + sizeof(double) * 2 * HB_CHUNK * HB_BLOCK) + 2 * warp;
        br.buf = stage + lane;
        br.sbuf = (uint32_t)__cvta_generic_to_shared(stage);
        br.mbar[0] = (uint32_t)__cvta_generic_to_shared(bars);
        br.mbar[1] = br.mbar[0] + 8;
        br.wmask = __activemask();
        br.leader = lane == 0;
        if (br.leader) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(br.mbar[0]) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(br.mbar[1]) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp(br.wmask);
    }
    norm3_fast_chain<VIT, FB, LL, BULK, TL>(a, seg, b, m, br, live);
}

template <bool VIT, bool FB>
__global__ void __launch_bounds__(HB_BLOCK)
    binom2_kernel(const __grid_constant__ Binom2Args a)
{
    extern __shared__ double smem[];
    const int32_t seg = a.seg_order[blockIdx.x / a.nblk_per_seg];
    const int64_t b = (int64_t)(blockIdx.x % a.nblk_per_seg) * HB_BLOCK + threadIdx.x;
    if (b >= a.B) return;
    if (a.gate_in && *a.gate_in == 0) return; // the lean kernel handled this launch
    binom2_chain<VIT, FB>(a, seg, b, smem + threadIdx.x, HB_BLOCK);
}

// Emission pre-pass for pooled chains: one thread per (position, chain).
__global__ void __launch_bounds__(128)
    binom2_emit_kernel(const __grid_constant__ Binom2Args a, double2 *pre_lb)
{
    const int64_t b = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t t = blockIdx.x; // positions on grid.x (can exceed 65535)
    if (b >= a.B || t >= a.T) return;
    const int64_t o = t * a.ld + b;
    double l0 = 0, l1 = 0;
    RhoME rho;
    rho.m = 1.0;
    rho.e = 0;
    binom2_emit(a, a.x[o], a.n[o], l0, l1, rho, true, false);
    double2 v;
    v.x = l0;
    v.y = l1;
    pre_lb[o] = v;
}

#ifndef HB_B2F_MINB
#define HB_B2F_MINB 4
#endif
template <bool VIT, bool FB, bool LL, bool PK, bool TL>
__global__ void __launch_bounds__(HB_BLOCK, HB_B2F_MINB)
    binom2_fast_kernel(const __grid_constant__ Binom2Args a)
{
    extern __shared__ double2 smem2[];
    const int32_t seg = a.seg_order[blockIdx.x / a.nblk_per_seg];
    const int64_t b = (int64_t)(blockIdx.x % a.nblk_per_seg) * HB_BLOCK + threadIdx.x;
    // block-local copy of the emission table for small trial counts (before any thread may leave)
    double2 *slb = smem2;
    double *srho = reinterpret_cast<double *>(smem2 + HB_B2F_SLUT_ENTRIES);
    const int32_t slut_n = a.lut_nmax < HB_B2F_SLUT_N ? a.lut_nmax : HB_B2F_SLUT_N;
    const int32_t nent = (slut_n + 1) * (slut_n + 2) / 2;
    for (int32_t i = threadIdx.x; i < nent; i += HB_BLOCK) {
        slb[i] = a.lut_lb[i];
        srho[i] = a.lut_rho_d[i];
    }
    __syncthreads();
    int tix = (int)threadIdx.x;
    int64_t bb = b;
    if (TL) { // complete warps for the cooperative staging: lanes past the last chain repeat it
        if ((b & ~(int64_t)31) >= a.B) return;
        if (b >= a.B) {
            tix -= (int)(b - (a.B - 1));
            bb = a.B - 1;
        }
    } else if (b >= a.B) {
        return;
    }
    double2 *park = reinterpret_cast<double2 *>(reinterpret_cast<char *>(smem2) + HB_B2F_SLUT_BYTES);
    B2fSmem m;
    m.st = HB_BLOCK;
    m.pr = park + tix;
    m.ph = reinterpret_cast<int32_t *>(park + HB_CHUNK * HB_BLOCK) + tix;
    m.xs = m.ph + HB_CHUNK * HB_BLOCK;
    m.ns = m.xs + HB_CHUNK * HB_BLOCK;
    m.slb = slb;
    m.srho = srho;
    m.slut_n = slut_n;
    binom2_fast_chain<VIT, FB, LL, PK, TL>(a, seg, bb, m);
}

template <bool SYM>
static cudaError_t launch_norm3_t(const Norm3Args &a, bool vit, bool fb, cudaStream_t st)
{
    const unsigned grid = (unsigned)((int64_t)a.nseg * a.nblk_per_seg);
    const size_t sm_fb = (size_t)HB_NORM3_SMEM_BYTES_PER_THREAD * HB_BLOCK;
    const bool ll = a.ll != nullptr;
    if (vit && fb) {
        if (ll)
            norm3_kernel<SYM, true, true, true><<<grid, HB_BLOCK, sm_fb, st>>>(a);
        else
            norm3_kernel<SYM, true, true, false><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    } else if (fb) {
        if (ll)
            norm3_kernel<SYM, false, true, true><<<grid, HB_BLOCK, sm_fb, st>>>(a);
        else
            norm3_kernel<SYM, false, true, false><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    } else {
        norm3_kernel<SYM, true, false, false>
            <<<grid, HB_BLOCK, (size_t)HB_CHUNK * HB_BLOCK * sizeof(double), st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_norm3(const Norm3Args &a, bool sym, bool vit, bool fb, cudaStream_t st)
{
    return sym ? launch_norm3_t<true>(a, vit, fb, st) : launch_norm3_t<false>(a, vit, fb, st);
}

#ifndef HB_FAST_BULK
// 1: on the tiled layout stage x with cp.async.bulk (TMA) + one mbarrier per stage and warp instead of
// the per-thread cp.async ring.  Works (bit-identical results) and removes ~13 instructions per step,
// but measured SLOWER on B200 (2.01 vs 1.95 ms on the probe workload): the kernel sits on the memory
// system's ceiling, not on its issue slots, and the per-chunk mbarrier wait + warp sync cost more than
// the ring's wait_group.  Kept as a build option (tools/build_variant.sh bulk -DHB_FAST_BULK=1).
#define HB_FAST_BULK 0
#endif
template <bool BULK, bool TL>
static cudaError_t launch_norm3_fast_t(const Norm3Args &a, bool vit, bool fb, cudaStream_t st)
{
    const unsigned grid = (unsigned)((int64_t)a.nseg * a.nblk_per_seg);
    const size_t sm_fb = BULK ? HB_FAST_BULK_SMEM_BYTES(HB_BLOCK) : (size_t)HB_FAST_SMEM_BYTES_PER_THREAD * HB_BLOCK;
    // Viterbi only: x ring (one chunk) + the warps' state tiles; the bulk form keeps the full carve-up
    const size_t sm_v = BULK ? sm_fb : (size_t)HB_CHUNK * HB_BLOCK * (sizeof(double) + 1);
    // > 48 KB of dynamic shared memory: opt in (per device, cheap)
    const bool ll = a.ll != nullptr;
#define HB_LAUNCH_FAST(V, F, L_, SM)                                                               \
    do {                                                                                          \
        cudaFuncSetAttribute(norm3_fast_kernel<V, F, L_, BULK, TL>,                                \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM));             \
        norm3_fast_kernel<V, F, L_, BULK, TL><<<grid, HB_BLOCK, (SM), st>>>(a);                    \
    } while (0)
    if (vit && fb) {
        if (ll) HB_LAUNCH_FAST(true, true, true, sm_fb);
        else HB_LAUNCH_FAST(true, true, false, sm_fb);
    } else if (fb) {
        if (ll) HB_LAUNCH_FAST(false, true, true, sm_fb);
        else HB_LAUNCH_FAST(false, true, false, sm_fb);
    } else {
        HB_LAUNCH_FAST(true, false, false, sm_v);
    }
#undef HB_LAUNCH_FAST
    return cudaGetLastError();
}

cudaError_t launch_norm3_fast(const Norm3Args &a, bool vit, bool fb, cudaStream_t st)
{
#if HB_FAST_BULK
    if (a.tiled) return launch_norm3_fast_t<true, true>(a, vit, fb, st);
#endif
    if (a.tiled) return launch_norm3_fast_t<false, true>(a, vit, fb, st);
    return launch_norm3_fast_t<false, false>(a, vit, fb, st);
}

// Chain-resident posteriors (hb_norm3_resident.cuh): one chain per block of NW warps, one launch per segment.
template <int NW>
__global__ void __launch_bounds__(32 * NW) norm3_resident_kernel(const __grid_constant__ FbrArgs a)
{
    extern __shared__ double fbr_smem[];
    fbr_chain<NW>(a, (int64_t)blockIdx.x, fbr_smem);
}

cudaError_t launch_norm3_resident(const FbrArgs &a, int nw, size_t smem, cudaStream_t st)
{
    const unsigned grid = (unsigned)a.B;
#define HB_LAUNCH_FBR(NW_)                                                                                    \
    do {                                                                                                      \
        cudaFuncSetAttribute(norm3_resident_kernel<NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
        norm3_resident_kernel<NW_><<<grid, 32 * NW_, smem, st>>>(a);                                          \
    } while (0)
    if (nw == 1) HB_LAUNCH_FBR(1);
    else if (nw == 2) HB_LAUNCH_FBR(2);
    else HB_LAUNCH_FBR(4);
#undef HB_LAUNCH_FBR
    return cudaGetLastError();
}

cudaError_t launch_binom2_emit(const Binom2Args &a, double2 *pre_lb, cudaStream_t st)
{
    dim3 grid((unsigned)a.T, (unsigned)((a.B + 127) / 128));
    binom2_emit_kernel<<<grid, 128, 0, st>>>(a, pre_lb);
    return cudaGetLastError();
}

template <bool PK, bool TL>
static cudaError_t launch_binom2_fast_t(const Binom2Args &a, bool vit, bool fb, cudaStream_t st)
{
    const unsigned grid = (unsigned)((int64_t)a.nseg * a.nblk_per_seg);
    const size_t sm_fb = (size_t)HB_B2F_SLUT_BYTES + (size_t)HB_B2F_SMEM_BYTES_PER_THREAD * HB_BLOCK;
    const bool ll = a.ll != nullptr;
    if (vit && fb) {
        if (ll)
            binom2_fast_kernel<true, true, true, PK, TL><<<grid, HB_BLOCK, sm_fb, st>>>(a);
        else
            binom2_fast_kernel<true, true, false, PK, TL><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    } else if (fb) {
        if (ll)
            binom2_fast_kernel<false, true, true, PK, TL><<<grid, HB_BLOCK, sm_fb, st>>>(a);
        else
            binom2_fast_kernel<false, true, false, PK, TL><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    } else {
        binom2_fast_kernel<true, false, false, PK, TL><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_binom2_fast(const Binom2Args &a, bool vit, bool fb, cudaStream_t st)
{
    if (a.tiled)
        return a.packed ? launch_binom2_fast_t<true, true>(a, vit, fb, st)
                        : launch_binom2_fast_t<false, true>(a, vit, fb, st);
    return a.packed ? launch_binom2_fast_t<true, false>(a, vit, fb, st)
                    : launch_binom2_fast_t<false, false>(a, vit, fb, st);
}

// (n << 16) | x per element; *overflow is set when a count does not fit 16 bits
__global__ void pack_counts_kernel(const int32_t *x, const int32_t *n, int64_t count, uint32_t *out,
                                   int32_t *overflow)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int32_t xv = x[i], nv = n[i];
    if ((uint32_t)xv > 65535u || (uint32_t)nv > 65535u) *overflow = 1;
    out[i] = ((uint32_t)nv << 16) | ((uint32_t)xv & 0xffffu);
}

cudaError_t launch_pack_counts(const int32_t *x, const int32_t *n, int64_t count, uint32_t *out,
                               int32_t *overflow, cudaStream_t st)
{
    pack_counts_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(x, n, count, out, overflow);
    return cudaGetLastError();
}

cudaError_t launch_binom2(const Binom2Args &a, bool vit, bool fb, cudaStream_t st)
{
    const unsigned grid = (unsigned)((int64_t)a.nseg * a.nblk_per_seg);
    const size_t sm_fb = (size_t)HB_CHUNK * HB_BINOM2_SLOT * HB_BLOCK * sizeof(double);
    if (vit && fb)
        binom2_kernel<true, true><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    else if (fb)
        binom2_kernel<false, true><<<grid, HB_BLOCK, sm_fb, st>>>(a);
    else
        binom2_kernel<true, false><<<grid, HB_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}

// =====================================================================================
// Few long chains: Viterbi parallel along the gene axis (hb_longchain.cuh)
// =====================================================================================
#define LC_TPB 32 // threads per block of the parallel phases: spread the few thousand walkers over all SMs
template <int S, int PH>
__global__ void __launch_bounds__(LC_TPB) lc_par_kernel(const __grid_constant__ LcArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * LC_TPB + threadIdx.x;
    if (i >= (int64_t)a.K * a.nb) return;
    const int32_t c = (int32_t)(i / a.nb);
    const int64_t b = i % a.nb;
    if (PH == 0) lc_fmat<S>(a, c, b);
    if (PH == 2) lc_imat<S>(a, c, b);
    if (PH == 4) lc_literal<S>(a, c, b);
    if (PH == 6) lc_write_y<S>(a, c, b);
}
// Sequential walkers: one warp per chain.  The warp stages the per-block tables of HB_LC_WIN blocks in
// shared memory (coalesced), lane 0 walks them (a dependent recursion over the blocks).
template <typename T>
__device__ __forceinline__ void lc_stage(T *dst, const T *src, int64_t count, int lane)
{
    for (int64_t i = lane; i < count; i += 32) dst[i] = src[i];
}

template <int S>
__global__ void __launch_bounds__(32) lc_p1_kernel(const __grid_constant__ LcArgs a)
{
    __shared__ double s_f[HB_LC_WIN * S * S];
    __shared__ int32_t s_be[HB_LC_WIN];
    const int32_t c = blockIdx.x;
    const int lane = threadIdx.x;
    double v[S];
    if (lane == 0) lc_prefix_approx_begin<S>(a, c, v);
    for (int64_t w0 = 0; w0 < a.nb; w0 += HB_LC_WIN) {
        const int64_t n = a.nb - w0 < HB_LC_WIN ? a.nb - w0 : HB_LC_WIN;
        lc_stage(s_f, a.fmat + ((int64_t)c * a.nb + w0) * S * S, n * S * S, lane);
        __syncwarp();
        if (lane == 0) {
            LcArgs w = a;
            w.win_b0 = w0;
            w.wf = s_f;
            lc_prefix_approx_range<S>(w, c, w0, w0 + n, v, s_be);
        }
        __syncwarp();
        lc_stage(a.bexp + (int64_t)c * a.nb + w0, (const int32_t *)s_be, n, lane);
        __syncwarp();
    }
}

template <int S>
__global__ void __launch_bounds__(32) lc_p3_kernel(const __grid_constant__ LcArgs a, int32_t round)
{
    __shared__ int64_t s_i[HB_LC_WIN * 2 * S * S];
    __shared__ int32_t s_be[HB_LC_WIN];
    __shared__ uint8_t s_io[HB_LC_WIN];
    const int32_t c = blockIdx.x;
    const int lane = threadIdx.x;
    double nu[S];
    int64_t b0 = 0, forced_end = 0;
    int32_t nlit = 0;
    bool active = false;
    if (lane == 0) active = lc_prefix_exact_begin<S>(a, c, round, nu, b0, forced_end);
    active = __shfl_sync(0xffffffffu, active ? 1 : 0, 0) != 0;
    if (!active) return;
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    for (int64_t w0 = b0; w0 < a.nb; w0 += HB_LC_WIN) {
        const int64_t n = a.nb - w0 < HB_LC_WIN ? a.nb - w0 : HB_LC_WIN;
        const int64_t o = (int64_t)c * a.nb + w0;
        lc_stage(s_i, (const int64_t *)a.imat + o * 2 * S * S, n * 2 * S * S, lane);
        lc_stage(s_be, (const int32_t *)a.bexp + o, n, lane);
        lc_stage(s_io, (const uint8_t *)a.iok + o, n, lane);
        __syncwarp();
        if (lane == 0) {
            LcArgs w = a;
            w.win_b0 = w0;
            w.wi = s_i;
            w.wbe = s_be;
            w.wio = s_io;
            lc_prefix_exact_range<S>(w, c, w0, w0 + n, nu, forced_end, nlit);
        }
        __syncwarp();
    }
    if (lane == 0) a.cstate[c * HB_LC_CS + 3] += nlit;
}

template <int S>
__global__ void __launch_bounds__(32) lc_p5_kernel(const __grid_constant__ LcArgs a)
{
    if (threadIdx.x == 0) lc_resolve<S>(a, blockIdx.x);
}

template <int S>
static cudaError_t launch_longchain_t(const LcArgs &a, cudaStream_t st)
{
    const int64_t nw = (int64_t)a.K * a.nb;
    const unsigned gp = (unsigned)((nw + LC_TPB - 1) / LC_TPB);
    if (nw > 0) {
        lc_par_kernel<S, 0><<<gp, LC_TPB, 0, st>>>(a);
        lc_p1_kernel<S><<<a.K, 32, 0, st>>>(a);
        lc_par_kernel<S, 2><<<gp, LC_TPB, 0, st>>>(a);
        for (int r = 0; r < HB_LC_ROUNDS; r++) {
            lc_p3_kernel<S><<<a.K, 32, 0, st>>>(a, r);
            lc_par_kernel<S, 4><<<gp, LC_TPB, 0, st>>>(a);
        }
    }
    lc_p5_kernel<S><<<a.K, 32, 0, st>>>(a);
    if (nw > 0) lc_par_kernel<S, 6><<<gp, LC_TPB, 0, st>>>(a);
    return cudaGetLastError();
}

// ---- posteriors of the same chains (lc_fb_*)
template <int S, int PH>
__global__ void __launch_bounds__(LC_TPB) lc_fb_par_kernel(const __grid_constant__ LcArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * LC_TPB + threadIdx.x;
    if (i >= (int64_t)a.K * a.nb) return;
    const int32_t c = (int32_t)(i / a.nb);
    const int64_t b = i % a.nb;
    if (PH == 0) lc_fb_mat<S>(a, c, b);
    if (PH == 2) lc_fb_block<S>(a, c, b);
}
template <int S>
__global__ void __launch_bounds__(32) lc_fb_p1_kernel(const __grid_constant__ LcArgs a)
{
    __shared__ double s_f[HB_LC_WIN * S * S];
    __shared__ int32_t s_e[HB_LC_WIN * S];
    const int32_t c = blockIdx.x;
    const int lane = threadIdx.x;
    double al[S], bt[S];
    if (lane == 0) {
        lc_fb_begin<S>(a, c, al);
#pragma unroll
        for (int k = 0; k < S; k++) bt[k] = 1.0 / S;
    }
    for (int dir = 0; dir < 2; dir++) {
        const int64_t nwin = (a.nb + HB_LC_WIN - 1) / HB_LC_WIN;
        for (int64_t wi = 0; wi < nwin; wi++) {
            // forward: windows from the front; backward: from the back, aligned to the same cuts
            const int64_t w0 = (dir == 0 ? wi : nwin - 1 - wi) * HB_LC_WIN;
            const int64_t n = a.nb - w0 < HB_LC_WIN ? a.nb - w0 : HB_LC_WIN;
            const int64_t o = (int64_t)c * a.nb + w0;
            lc_stage(s_f, (const double *)a.ffm + o * S * S, n * S * S, lane);
            lc_stage(s_e, (const int32_t *)a.ffe + o * S, n * S, lane);
            __syncwarp();
            if (lane == 0) {
                LcArgs w = a;
                w.win_b0 = w0;
                w.wff = s_f;
                w.wfe = s_e;
                if (dir == 0)
                    lc_fb_fwd_range<S>(w, c, w0, w0 + n, al);
                else
                    lc_fb_bwd_range<S>(w, c, w0, w0 + n, bt);
            }
            __syncwarp();
        }
    }
}
template <int S>
__global__ void __launch_bounds__(32) lc_fb_p3_kernel(const __grid_constant__ LcArgs a)
{
    if (threadIdx.x == 0) lc_fb_ll<S>(a, blockIdx.x);
}
template <int S>
static cudaError_t launch_longchain_fb_t(const LcArgs &a, cudaStream_t st)
{
    const int64_t nw = (int64_t)a.K * a.nb;
    const unsigned gp = (unsigned)((nw + LC_TPB - 1) / LC_TPB);
    if (nw > 0) {
        lc_fb_par_kernel<S, 0><<<gp, LC_TPB, 0, st>>>(a);
        lc_fb_p1_kernel<S><<<a.K, 32, 0, st>>>(a);
        lc_fb_par_kernel<S, 2><<<gp, LC_TPB, 0, st>>>(a);
    }
    lc_fb_p3_kernel<S><<<a.K, 32, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_longchain_fb(const LcArgs &a, cudaStream_t st)
{
    return a.S == 2 ? launch_longchain_fb_t<2>(a, st) : launch_longchain_fb_t<3>(a, st);
}

cudaError_t launch_longchain(const LcArgs &a, cudaStream_t st)
{
    return a.S == 2 ? launch_longchain_t<2>(a, st) : launch_longchain_t<3>(a, st);
}

// Literal log-emissions of pooled chains, chain-major [K][G][S]: one thread per (chain, position).
__global__ void __launch_bounds__(128)
    binom2_emit_cm_kernel(const __grid_constant__ Binom2Args a, const int32_t *x, const int32_t *n,
                          int64_t count, double2 *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double l0 = 0, l1 = 0;
    RhoME rho;
    rho.m = 1.0;
    rho.e = 0;
    binom2_emit(a, x[i], n[i], l0, l1, rho, true, false);
    double2 v;
    v.x = l0;
    v.y = l1;
    out[i] = v;
}
cudaError_t launch_binom2_emit_cm(const Binom2Args &a, const int32_t *x, const int32_t *n, int64_t count,
                                  double2 *out, cudaStream_t st)
{
    binom2_emit_cm_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(a, x, n, count, out);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
    norm3_emit_cm_kernel(const __grid_constant__ Norm3Const c, const double *x, int64_t G, int32_t K,
                         const double *sd, const double *inv_sd, const double *log_sd, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * K) return;
    const int64_t k = i / G;
    double l0, l1, l2, h1;
    norm3_logb(x[i], c, sd[k], inv_sd[k], log_sd[k], l0, l1, l2, h1);
    out[3 * i] = l0;
    out[3 * i + 1] = l1;
    out[3 * i + 2] = l2;
}
cudaError_t launch_norm3_emit_cm(const Norm3Const &c, const double *x, int64_t G, int32_t K, const double *sd,
                                 const double *inv_sd, const double *log_sd, double *out, cudaStream_t st)
{
    norm3_emit_cm_kernel<<<(unsigned)((G * K + 127) / 128), 128, 0, st>>>(c, x, G, K, sd, inv_sd, log_sd, out);
    return cudaGetLastError();
}

// =====================================================================================
// Per-chain emission scale helpers
// =====================================================================================
__global__ void sd_prep_kernel(const double *sd, int64_t n, double *inv_sd, double *log_sd)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double s = sd[i];
        inv_sd[i] = 1.0 / s; // IEEE division: the correctly rounded reciprocal div_by_const needs
        log_sd[i] = log(s);
    }
}

cudaError_t launch_sd_prep(const double *sd, int64_t n, double *inv_sd, double *log_sd,
                           cudaStream_t st)
{
    sd_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sd, n, inv_sd, log_sd);
    return cudaGetLastError();
}

// sd over positions of every (column, segment) chain: two passes, compensated sums (TwoSum).
__device__ __forceinline__ void two_sum_acc(double &s, double &c, double v)
{
    double t = s + v;
    double bp = t - s;
    c += (s - (t - bp)) + (v - bp);
    s = t;
}

__global__ void chain_sd_kernel(const double *x, int64_t ld, int64_t B, const int64_t *seg_off,
                                int32_t nseg, double *sd)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t seg = blockIdx.y;
    if (b >= B || seg >= nseg) return;
    const int64_t t0 = seg_off[seg], t1 = seg_off[seg + 1];
    const int64_t L = t1 - t0;
    if (L < 2) {
        sd[(int64_t)seg * B + b] = NAN;
        return;
    }
    double s = 0, c = 0;
    for (int64_t t = t0; t < t1; t++) two_sum_acc(s, c, x[t * ld + b]);
    const double mean = (s + c) / (double)L;
    double q = 0, qc = 0;
    for (int64_t t = t0; t < t1; t++) {
        double d = x[t * ld + b] - mean;
        double p = d * d;
        double pe = __fma_rn(d, d, -p);
        two_sum_acc(q, qc, p);
        qc += pe;
    }
    sd[(int64_t)seg * B + b] = sqrt((q + qc) / (double)(L - 1));
}

cudaError_t launch_chain_sd(const double *x, int64_t ld, int64_t B, const int64_t *seg_off_dev,
                            int32_t nseg, double *sd, cudaStream_t st)
{
    dim3 grid((unsigned)((B + 127) / 128), (unsigned)nseg);
    chain_sd_kernel<<<grid, 128, 0, st>>>(x, ld, B, seg_off_dev, nseg, sd);
    return cudaGetLastError();
}

// =====================================================================================
// Layout: R column-major [G x N] (src[c*G + g]) <-> gene-major dst[g*ld + c]
// =====================================================================================
template <typename TI, typename TO, bool RINT>
__global__ void transpose_in_kernel(const TI *src, int64_t G, int64_t N, TO *dst, int64_t ld)
{
    __shared__ TI tile[32][33];
    const int64_t g0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t c = c0 + r, g = g0 + threadIdx.x;
        if (c < N && g < G) tile[r][threadIdx.x] = src[c * G + g];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t g = g0 + r, c = c0 + threadIdx.x;
        if (g < G && c < N) {
            TI v = tile[threadIdx.x][r];
            dst[g * ld + c] = RINT ? (TO)rint((double)v) : (TO)v;
        }
    }
}

// gene-major src[g*ld + c] -> column-major dst[c*G + g]
template <typename TI, typename TO>
__global__ void transpose_out_kernel(const TI *src, int64_t ld, int64_t G, int64_t N, TO *dst)
{
    __shared__ TI tile[32][33];
    const int64_t g0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t g = g0 + r, c = c0 + threadIdx.x;
        if (g < G && c < N) tile[r][threadIdx.x] = src[g * ld + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t c = c0 + r, g = g0 + threadIdx.x;
        if (c < N && g < G) dst[c * G + g] = (TO)tile[threadIdx.x][r];
    }
}

static dim3 tgrid(int64_t G, int64_t N) { return dim3((unsigned)((G + 31) / 32), (unsigned)((N + 31) / 32)); }

cudaError_t launch_transpose_f64(const double *src, int64_t G, int64_t N, double *dst, int64_t ld,
                                 cudaStream_t st)
{
    transpose_in_kernel<double, double, false><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, G, N, dst, ld);
    return cudaGetLastError();
}
cudaError_t launch_transpose_f64_i32(const double *src, int64_t G, int64_t N, int32_t *dst,
                                     int64_t ld, cudaStream_t st)
{
    transpose_in_kernel<double, int32_t, true><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, G, N, dst, ld);
    return cudaGetLastError();
}
cudaError_t launch_transpose_i32(const int32_t *src, int64_t G, int64_t N, int32_t *dst,
                                 int64_t ld, cudaStream_t st)
{
    transpose_in_kernel<int32_t, int32_t, false><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, G, N, dst, ld);
    return cudaGetLastError();
}
cudaError_t launch_transpose_u8_back(const uint8_t *src, int64_t ld, int64_t G, int64_t N,
                                     int32_t *dst, cudaStream_t st)
{
    transpose_out_kernel<uint8_t, int32_t><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, ld, G, N, dst);
    return cudaGetLastError();
}
cudaError_t launch_transpose_u8_back_u8(const uint8_t *src, int64_t ld, int64_t G, int64_t N,
                                        uint8_t *dst, cudaStream_t st)
{
    transpose_out_kernel<uint8_t, uint8_t><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, ld, G, N, dst);
    return cudaGetLastError();
}
cudaError_t launch_transpose_f64_back(const double *src, int64_t ld, int64_t G, int64_t N,
                                      double *dst, cudaStream_t st)
{
    transpose_out_kernel<double, double><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, ld, G, N, dst);
    return cudaGetLastError();
}

cudaError_t launch_transpose_i32_back(const int32_t *src, int64_t ld, int64_t G, int64_t N,
                                      int32_t *dst, cudaStream_t st)
{
    transpose_out_kernel<int32_t, int32_t><<<tgrid(G, N), dim3(32, 8), 0, st>>>(src, ld, G, N, dst);
    return cudaGetLastError();
}

// gene-major src[t*ld + b] <-> warp-tiled dst[((b/32)*T + t)*32 + b%32]; both sides coalesced
// (a warp handles one 32-column tile row).  Padding columns of the last tile are written as 0.
template <typename E>
__global__ void tile_kernel(const E *src, int64_t ld, int64_t T, int64_t B, E *dst, bool to_tiled)
{
    const int64_t Bp = (B + 31) / 32 * 32;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bp) return;
    for (int64_t t = blockIdx.y; t < T; t += gridDim.y) {
        const int64_t ot = (((b >> 5) * T + t) << 5) + (b & 31);
        if (to_tiled)
            dst[ot] = b < B ? src[t * ld + b] : E(0);
        else if (b < B)
            dst[t * ld + b] = src[ot];
    }
}

cudaError_t launch_tile(const void *src, int64_t ld, int64_t T, int64_t B, int elem, void *dst,
                        bool to_tiled, cudaStream_t st)
{
    const int64_t Bp = (B + 31) / 32 * 32;
    dim3 grid((unsigned)((Bp + 255) / 256), (unsigned)std::min<int64_t>(T, 65535));
    if (elem == 1)
        tile_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src, ld, T, B, (uint8_t *)dst, to_tiled);
    else if (elem == 4)
        tile_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)src, ld, T, B, (uint32_t *)dst, to_tiled);
    else
        tile_kernel<uint64_t><<<grid, 256, 0, st>>>((const uint64_t *)src, ld, T, B, (uint64_t *)dst, to_tiled);
    return cudaGetLastError();
}

// =====================================================================================
// Deterministic retest of a candidate region, expression model: the exact
// marginal posterior of inst/bug/expressionModel.bug, which the reference samples with JAGS
// (HoneyBADGER$calcGexpCnvProb, R/HoneyBADGER.R:374-470, called from :1220).  Priors integrate to
// P(S_k = 1) = 1/2 per cell and P(dd = 1) = 1/2 shared; with l_k(mu) = sum_j log N(gexp[j,k]; mu,
// precision sigma0_k):  L_k(d) = (exp l_k(0) + exp l_k(+-mag0)) / 2,  P(dd = 1 | data) =
// sigmoid(sum_k log L_k(1) - log L_k(0)),  w_k(d) = exp l_k(+-mag0) / (2 L_k(d)),
// P(amp_k) = P(dd=1) w_k(1),  P(del_k) = P(dd=0) w_k(0)  (= mean(S dd), mean(S (1-dd)), :460-461).
// Phase A: one thread per cell streams the region's rows (coalesced along cells), writes w_k and
// the block partials of D = sum_k log L_k(1) - log L_k(0).  Phase B: every block re-sums the
// partials in a fixed order (deterministic), applies the sigmoid and scales.
// Only differences against l_k(0) enter: dp = l_k(+mag0) - l_k(0) = prec (mag0 s1 - n mag0^2 / 2),
// dm likewise with -mag0;  w_k(1) = sigmoid(dp), w_k(0) = sigmoid(dm),
// log L_k(1) - log L_k(0) = softplus(dp) - softplus(dm)   (s1 = sum_j gexp[j,k]).
__device__ __forceinline__ double softplus(double d) { return d > 0 ? d + log1p(exp(-d)) : log1p(exp(d)); }
__device__ __forceinline__ double sigmoid(double d)
{
    return d >= 0 ? 1.0 / (1.0 + exp(-d)) : exp(d) / (1.0 + exp(d));
}

__global__ void __launch_bounds__(128)
    retest_gexp_a_kernel(const double *x, int64_t ld, int64_t N, const int64_t *rows, int64_t nrows,
                         double mag0, const double *sigma0, const double *del_off, double *w_amp,
                         double *w_del, double *mu0, double *partial)
{
    __shared__ double red[128];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double dk = 0.0;
    if (c < N) {
        double s1 = 0.0;
        for (int64_t j = 0; j < nrows; j++) s1 += ld_stream(x + rows[j] * ld + c);
        const double n = (double)nrows, prec = sigma0 ? sigma0[c] : 1.0;
        const double half = 0.5 * n * mag0 * mag0;
        // combined model (inst/bug/combinedModel.bug): the deletion hypothesis also switches the cell's
        // allele likelihood from the neutral to the deleted form — a per-cell log-odds added to dm
        const double dp = prec * (mag0 * s1 - half), dm = prec * (-mag0 * s1 - half) + (del_off ? del_off[c] : 0.0);
        w_amp[c] = sigmoid(dp);
        w_del[c] = sigmoid(dm);
        if (mu0) mu0[c] = s1 / n;
        dk = softplus(dp) - softplus(dm);
    }
    red[threadIdx.x] = dk;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// D_total: sum of the partials of ALL ranks (device scalar), or NULL to sum this launch's partials
__global__ void __launch_bounds__(128)
    retest_gexp_b_kernel(int64_t N, const double *partial, int32_t npartial, const double *D_total,
                         double *p_amp, double *p_del)
{
    __shared__ double Dsh;
    if (threadIdx.x == 0) {
        double D = 0.0;
        if (D_total)
            D = *D_total;
        else
            for (int32_t i = 0; i < npartial; i++) D += partial[i];
        Dsh = D;
    }
    __syncthreads();
    const double post1 = sigmoid(Dsh);
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < N) {
        p_amp[c] *= post1;
        p_del[c] *= (1.0 - post1);
    }
}

__global__ void sum_partials_kernel(const double *partial, int32_t n, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double D = 0.0;
        for (int32_t i = 0; i < n; i++) D += partial[i];
        *out = D;
    }
}

// Allele region retest, likelihood-ratio stand-in for calcAlleleCnvProb's JAGS snpModel
// (R/HoneyBADGER.R:907-1042): P(deletion) per cell = sigmoid(sum over the region's SNPs of
// log dbinom(r, n, pd) - log dbinom(r, n, pn)) with equal priors; the binomial coefficient cancels.
__global__ void __launch_bounds__(128)
    retest_allele_kernel(const int32_t *r, const int32_t *n, int64_t ld, int64_t N, const int64_t *rows,
                         int64_t nrows, double lr_hit, double lr_miss, double *p_del)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    int64_t sr = 0, sn = 0; // the log-ratio is linear in the counts: integer sums, exact
    for (int64_t j = 0; j < nrows; j++) {
        const int64_t o = rows[j] * ld + c;
        sr += ld_stream(r + o);
        sn += ld_stream(n + o);
    }
    p_del[c] = sigmoid((double)sr * lr_hit + (double)(sn - sr) * lr_miss);
}

cudaError_t launch_retest_allele(const int32_t *r, const int32_t *n, int64_t ld, int64_t N,
                                 const int64_t *rows, int64_t nrows, double pd, double pn, double *p_del,
                                 cudaStream_t st)
{
    retest_allele_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(
        r, n, ld, N, rows, nrows, log(pd) - log(pn), log(1 - pd) - log(1 - pn), p_del);
    return cudaGetLastError();
}

// Allele region retest, the model itself: exact marginal posterior P(S_k = 1 | data) of
// inst/bug/snpModel.bug (sampled with JAGS by calcAlleleCnvProb, R/HoneyBADGER.R:907-1042) with every
// per-cell latent variable integrated out —
//   S_k ~ Bern(alpha_k), alpha_k ~ U(0,1)                      => prior odds 1
//   S_k = 1:  r ~ Bin(n, fma),  fma = pseudo (ma = 1) | 1 - pseudo (ma = 0)
//   S_k = 0:  per gene j: b ~ Bern(mono), d ~ Bern(1/2);  b = 1: every SNP of the gene has p = pseudo (d = 1)
//             or 1 - pseudo (d = 0);  b = 0: p = h, h ~ N(0.5, precision 0.1) truncated to [pseudo, 1-pseudo],
//             independent per SNP: the integral over h is a 256-node composite Gauss-Legendre rule
// — and the per-SNP allele indicator ma, which the cells share through the bulk, replaced by its bulk
// posterior q = P(ma = 1 | l, n.bulk) (the one approximation; the bulk counts dominate it).
// One thread per cell; the binomial coefficients are common to both hypotheses and dropped.
// tab = [lh | l1h | lw][nq]: log h_q, log(1 - h_q), log(weight_q * density(h_q)); rows are gene-grouped,
// gene_end[j] marks the last SNP of a gene; lq1 / lq0 = log q, log(1 - q) per SNP.
#define RA_THREADS 128
__device__ __forceinline__ double lse2(double a, double b)
{
    const double m = fmax(a, b);
    if (m == -INFINITY) return m;
    return m + log(exp(a - m) + exp(b - m));
}
__global__ void __launch_bounds__(RA_THREADS)
    retest_allele_model_kernel(const int32_t *r, const int32_t *n, int64_t ld, int64_t N, const int64_t *rows,
                               int64_t nrows, const uint8_t *gene_end, const double *lq1, const double *lq0,
                               const double *tab, int32_t nq, double lps, double l1ps, double lmono_half,
                               double l1mono, int32_t logodds, double *p_del)
{
    extern __shared__ double sh[];
    for (int i = threadIdx.x; i < 3 * nq; i += RA_THREADS) sh[i] = tab[i];
    __syncthreads();
    const double *lh = sh, *l1h = sh + nq, *lw = sh + 2 * nq;
    const int64_t c = (int64_t)blockIdx.x * RA_THREADS + threadIdx.x;
    if (c >= N) return;
    double L1 = 0.0, L0 = 0.0, A = 0.0, B1 = 0.0, B0 = 0.0;
    for (int64_t j = 0; j < nrows; j++) {
        const int64_t o = rows[j] * ld + c;
        const int32_t nv = ld_stream(n + o);
        if (nv > 0) {
            const double rv = (double)ld_stream(r + o), mv = (double)nv - rv;
            const double a = rv * lps + mv * l1ps, b = rv * l1ps + mv * lps;
            L1 += lse2(lq1[j] + a, lq0[j] + b);
            B1 += a;
            B0 += b;
            double mx = -INFINITY;
            for (int q = 0; q < nq; q++) mx = fmax(mx, lw[q] + rv * lh[q] + mv * l1h[q]);
            double sum = 0.0;
            for (int q = 0; q < nq; q++) sum += exp(lw[q] + rv * lh[q] + mv * l1h[q] - mx);
            A += mx + log(sum);
        }
        if (gene_end[j]) {
            L0 += lse2(l1mono + A, lmono_half + lse2(B1, B0));
            A = B1 = B0 = 0.0;
        }
    }
    p_del[c] = logodds ? L1 - L0 : sigmoid(L1 - L0);
}

cudaError_t launch_retest_allele_model(const int32_t *r, const int32_t *n, int64_t ld, int64_t N,
                                       const int64_t *rows, int64_t nrows, const uint8_t *gene_end,
                                       const double *lq1, const double *lq0, const double *tab, int32_t nq,
                                       double pseudo, double mono, int32_t logodds, double *p_del,
                                       cudaStream_t st)
{
    retest_allele_model_kernel<<<(unsigned)((N + RA_THREADS - 1) / RA_THREADS), RA_THREADS,
                                 (size_t)3 * nq * sizeof(double), st>>>(
        r, n, ld, N, rows, nrows, gene_end, lq1, lq0, tab, nq, log(pseudo), log(1 - pseudo), log(mono / 2),
        log(1 - mono), logodds, p_del);
    return cudaGetLastError();
}

cudaError_t launch_retest_gexp_a(const double *x, int64_t ld, int64_t N, const int64_t *rows,
                                 int64_t nrows, double mag0, const double *sigma0, const double *del_off,
                                 double *w_amp, double *w_del, double *mu0, double *partial, double *D_local,
                                 cudaStream_t st)
{
    const unsigned nb = (unsigned)((N + 127) / 128);
    retest_gexp_a_kernel<<<nb, 128, 0, st>>>(x, ld, N, rows, nrows, mag0, sigma0, del_off, w_amp, w_del, mu0,
                                             partial);
    if (D_local) sum_partials_kernel<<<1, 32, 0, st>>>(partial, (int32_t)nb, D_local);
    return cudaGetLastError();
}
cudaError_t launch_retest_gexp_b(int64_t N, const double *partial, const double *D_total, double *p_amp,
                                 double *p_del, cudaStream_t st)
{
    const unsigned nb = (unsigned)((N + 127) / 128);
    retest_gexp_b_kernel<<<nb, 128, 0, st>>>(N, partial, (int32_t)nb, D_total, p_amp, p_del);
    return cudaGetLastError();
}

// =====================================================================================
// Grouping of cells (hb_group.cuh)
// tmp (ratio form only): G * ldo doubles of scratch for the ratios, or NULL (the ratio is then re-formed per window position)
cudaError_t launch_runmean(const double *x, const int32_t *r, const int32_t *n, int64_t ld, int64_t G,
                           int64_t N, int32_t k, double *out, int64_t ldo, double *tmp, cudaStream_t st)
{
    dim3 grid((unsigned)G, (unsigned)((N + 255) / 256));
    if (x)
        runmean_kernel<false><<<grid, 256, 0, st>>>(x, nullptr, nullptr, ld, G, N, k / 2, out, ldo);
    else if (tmp) {
        ratio_kernel<<<grid, 256, 0, st>>>(r, n, ld, G, N, tmp, ldo);
        runmean_kernel<false><<<grid, 256, 0, st>>>(tmp, nullptr, nullptr, ldo, G, N, k / 2, out, ldo);
    } else
        runmean_kernel<true><<<grid, 256, 0, st>>>(nullptr, r, n, ld, G, N, k / 2, out, ldo);
    return cudaGetLastError();
}
// scratch: dist_scratch_bytes(G, N) for the pipelined form (NULL, odd ld or a misaligned S: the plain kernel) —
// a byte per (slab, tile) and a 16-bit mask per (slab, cell)
size_t dist_scratch_bytes(int64_t G, int64_t N)
{
    const size_t nslab = (size_t)((G + GD_K - 1) / GD_K), nt = (size_t)((N + GD_T - 1) / GD_T);
    return ((nslab * nt + 255) & ~(size_t)255) + nslab * nt * GD_T * sizeof(uint16_t) + 256;
}
static cudaError_t dist_pipe_attr()
{
    constexpr int smem = GP_STAGES * 2 * GD_K * GD_T * (int)sizeof(double);
    static uint64_t done = 0; // function attributes are per device: one bit per device id (callers hold the API lock)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((done >> (dev & 63)) & 1)) {
        cudaError_t e = cudaFuncSetAttribute(dist_sq_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(dist_sq_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        done |= 1ull << (dev & 63);
    }
    return cudaSuccess;
}
static bool dist_pipe_ok(const double *S, int64_t ld, int64_t G, int64_t N)
{
    return N >= 2 && (ld & 1) == 0 && ((uintptr_t)S & 15) == 0 && (G + GD_K - 1) / GD_K <= 0x7fffffff &&
           (N + GD_T - 1) / GD_T <= 65535;
}
cudaError_t launch_dist_sq(const double *S, int64_t ld, int64_t G, int64_t N, double *D, void *scratch, cudaStream_t st)
{
    const unsigned nt = (unsigned)((N + GD_T - 1) / GD_T);
    const int64_t nslab = (G + GD_K - 1) / GD_K;
    if (scratch && dist_pipe_ok(S, ld, G, N)) {
        constexpr int smem = GP_STAGES * 2 * GD_K * GD_T * (int)sizeof(double);
        cudaError_t e = dist_pipe_attr();
        if (e != cudaSuccess) return e;
        uint8_t *flags = (uint8_t *)scratch;
        uint16_t *masks = (uint16_t *)((char *)scratch + (((size_t)nslab * nt + 255) & ~(size_t)255));
        dist_mask_kernel<<<dim3((unsigned)nslab, nt), GD_T, 0, st>>>(S, ld, G, N, masks, flags, (int)nt);
        dist_sq_pipe_kernel<true><<<dim3(nt, nt), 256, smem, st>>>(S, ld, N, masks, flags, (int)nt, S, ld, N, masks, flags,
                                                                   (int)nt, G, D, N);
        return cudaGetLastError();
    }
    dist_sq_kernel<<<dim3(nt, nt), 256, 0, st>>>(S, ld, G, N, D);
    return cudaGetLastError();
}
// Rectangular block: rows = the Na cells of A, columns = the Nb cells of B, D[i * ldd + j].  scratch:
// dist_scratch_bytes(G, Na) + dist_scratch_bytes(G, Nb).  Both matrices need an even stride and 16-byte alignment.
cudaError_t launch_dist_sq_rect(const double *Sa, int64_t lda, int64_t Na, const double *Sb, int64_t ldb, int64_t Nb,
                                int64_t G, double *D, int64_t ldd, void *scratch, cudaStream_t st)
{
    if (!scratch || !dist_pipe_ok(Sa, lda, G, Na) || !dist_pipe_ok(Sb, ldb, G, Nb)) return cudaErrorInvalidValue;
    constexpr int smem = GP_STAGES * 2 * GD_K * GD_T * (int)sizeof(double);
    cudaError_t e = dist_pipe_attr();
    if (e != cudaSuccess) return e;
    const unsigned nta = (unsigned)((Na + GD_T - 1) / GD_T), ntb = (unsigned)((Nb + GD_T - 1) / GD_T);
    const int64_t nslab = (G + GD_K - 1) / GD_K;
    uint8_t *fa = (uint8_t *)scratch;
    uint16_t *ma = (uint16_t *)((char *)scratch + (((size_t)nslab * nta + 255) & ~(size_t)255));
    char *sb = (char *)scratch + ((dist_scratch_bytes(G, Na) + 255) & ~(size_t)255);
    uint8_t *fb = (uint8_t *)sb;
    uint16_t *mb = (uint16_t *)(sb + (((size_t)nslab * ntb + 255) & ~(size_t)255));
    dist_mask_kernel<<<dim3((unsigned)nslab, nta), GD_T, 0, st>>>(Sa, lda, G, Na, ma, fa, (int)nta);
    dist_mask_kernel<<<dim3((unsigned)nslab, ntb), GD_T, 0, st>>>(Sb, ldb, G, Nb, mb, fb, (int)ntb);
    dist_sq_pipe_kernel<false><<<dim3(ntb, nta), 256, smem, st>>>(Sa, lda, Na, ma, fa, (int)nta, Sb, ldb, Nb, mb, fb, (int)ntb,
                                                                  G, D, ldd);
    return cudaGetLastError();
}
// Scratch of the linkage: the cluster form keeps one copy of the chain and the cluster sizes per CTA.
size_t ward_scratch_bytes(int64_t N) { return (size_t)N * sizeof(int32_t) * 2 * WC_MAXR + (size_t)N + 256; }

static cudaError_t ward_cluster_allow16()
{
    cudaError_t e = cudaSuccess;
    auto set = [&](auto k) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    };
    set(ward_cluster_kernel<1>), set(ward_cluster_kernel<2>), set(ward_cluster_kernel<3>), set(ward_cluster_kernel<4>);
    set(ward_cluster_kernel<6>), set(ward_cluster_kernel<8>), set(ward_cluster_kernel<WC_OWN>);
    return e;
}

// mode 0: cluster kernel from WARD_CLUSTER_MIN clusters up, single block below; 1: single block; 2: cluster
#define WARD_CLUSTER_MIN 256
cudaError_t launch_ward(double *D, int64_t N, void *work, int32_t *ma, int32_t *mb, double *height, int mode,
                        cudaStream_t st)
{
    int32_t *size = (int32_t *)work, *chain = size + (int64_t)N * WC_MAXR;
    if (mode != 1 && (mode == 2 || N >= WARD_CLUSTER_MIN)) {
        static int cluster_max_dev[64]; // per device: 0 not probed yet, else 1 + (16 non-portable size, 8, or 0 = none)
        int devno = 0;
        cudaGetDevice(&devno);
        int &slot = cluster_max_dev[devno & 63];
        int cluster_max = slot - 1;
        if (cluster_max < 0) {
            cluster_max = 0;
            for (int r : {16, 8}) {
                if (r > 8 && ward_cluster_allow16() != cudaSuccess) {
                    cudaGetLastError();
                    continue;
                }
                cudaLaunchConfig_t cfg = {};
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = r, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
                cfg.gridDim = dim3(r), cfg.blockDim = dim3(WC_THREADS), cfg.attrs = at, cfg.numAttrs = 1;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, ward_cluster_kernel<WC_OWN>, &cfg) == cudaSuccess && ncl >= 1) {
                    cluster_max = r;
                    break;
                }
                cudaGetLastError();
            }
            slot = cluster_max + 1;
        }
        // a CTA's slice must fit its threads' register masks: N / r + 32 <= WC_THREADS * WC_OWN (65 k cells at 16 CTAs)
        const bool fits = cluster_max && (N + cluster_max - 1) / cluster_max + 32 <= (int64_t)WC_THREADS * WC_OWN;
        if (fits) {
            // measured (tools/probe_ward.py): 8 CTAs are best up to ~8 k cells, 16 beyond; a step costs ~1.1 (1 k cells)
            // to ~2 us (10 k) — L2 latency + cluster barrier — against 1.3 to 6.5 us on one block, which wins only
            // below ~256 cells
            int r = std::min(cluster_max, N >= 8192 ? 16 : 8);
            if (const char *e = getenv("HB_WARD_R")) r = std::max(1, std::min(cluster_max, atoi(e))); // experiments
            if ((N + r - 1) / r + 32 > (int64_t)WC_THREADS * WC_OWN) r = cluster_max;
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = r, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(r), cfg.blockDim = dim3(WC_THREADS), cfg.stream = st, cfg.attrs = at, cfg.numAttrs = 1;
            const int64_t chunk = (((N + r - 1) / r) + 31) & ~(int64_t)31;
            const int own = (int)((chunk + WC_THREADS - 1) / WC_THREADS);
            if (own <= 1) return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<1>, D, N, size, chain, ma, mb, height);
            if (own <= 2) return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<2>, D, N, size, chain, ma, mb, height);
            if (own <= 3) return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<3>, D, N, size, chain, ma, mb, height);
            if (own <= 4) return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<4>, D, N, size, chain, ma, mb, height);
            if (own <= 6) return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<6>, D, N, size, chain, ma, mb, height);
            if (own <= 8) return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<8>, D, N, size, chain, ma, mb, height);
            return cudaLaunchKernelEx(&cfg, ward_cluster_kernel<WC_OWN>, D, N, size, chain, ma, mb, height);
        }
    }
    uint8_t *active = (uint8_t *)(chain + (int64_t)N * WC_MAXR);
    ward_nnchain_kernel<<<1, WL_THREADS, 0, st>>>(D, N, size, active, chain, ma, mb, height);
    return cudaGetLastError();
}

// =====================================================================================
// Pooled statistics of cell groups (R/HoneyBADGER.R:1141, :1149, :1404-1405)
// =====================================================================================
// Pooled expression observation, R `mean` semantics (summary.c real_mean: two passes in LDOUBLE, x87
// arithmetic replayed in integer registers, hb_x87.cuh).  The accumulation is order-dependent, so it
// is written as a RELAY over an 80-bit state per (group, gene): pool_accum_kernel continues `state`
// with this matrix's member cells in cell order — pass 0 adds x, pass 1 adds (x - mean80) — and the
// tiny kernels below close each pass.  One GPU runs accumulate / divide / accumulate / finish on
// the whole matrix; several GPUs hand `state` from rank to rank (honeybadger_b200/dist.py), which
// gives R's result bit for bit at any GPU count.
// Block = 32 genes x PM_KB groups; the gene tile is staged through shared memory in PM_TC-cell
// slabs so that global reads stay coalesced along the cell axis.
#define PM_TC 64
#define PM_KB 16
// One step of either pass on the double-pair form of the 80-bit state (hb_x87_pair.cuh): pass 0 adds v,
// pass 1 adds RN64(v - mean) with `negmean` = -mean.  Two roundings in pass 1, as R's loop has them.
__device__ __forceinline__ x80p pool_step(x80p S, x80p negmean, double v, bool centred)
{
    return centred ? x80p_add(S, x80p_add_double(negmean, v)) : x80p_add_double(S, v);
}
__global__ void __launch_bounds__(32 * PM_KB)
    pool_accum_kernel(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                      int32_t K, const x80 *mean80, x80 *state)
{
    __shared__ double tile[32][PM_TC + 1];
    const int64_t g0 = (int64_t)blockIdx.x * 32;
    const int lane = threadIdx.x, ky = threadIdx.y;
    const int tid = ky * 32 + lane;
    for (int32_t kb = 0; kb < K; kb += PM_KB) {
        const int32_t k = kb + ky;
        const bool live = k < K && g0 + lane < G;
        const uint8_t *mem = member + (int64_t)(live ? k : 0) * N;
        const int64_t si = (int64_t)(live ? k : 0) * G + (live ? g0 + lane : 0);
        x80p S = x80p_from_x80(live ? state[si] : x80_zero());
        const bool centred = mean80 != nullptr;
        const x80p M = x80p_from_x80((live && centred) ? x80_neg(mean80[si]) : x80_zero());
        for (int64_t c0 = 0; c0 < N; c0 += PM_TC) {
            __syncthreads();
            for (int i = tid; i < 32 * PM_TC; i += 32 * PM_KB) {
                int r = i / PM_TC, cc = i % PM_TC;
                int64_t g = g0 + r, c = c0 + cc;
                tile[r][cc] = (g < G && c < N) ? x[g * ld + c] : 0.0;
            }
            __syncthreads();
            if (live) {
                const int lim = (int)((N - c0) < PM_TC ? (N - c0) : PM_TC);
                for (int cc = 0; cc < lim; cc++) {
                    if (mem[c0 + cc]) S = pool_step(S, M, tile[lane][cc], centred);
                }
            }
        }
        if (live) state[si] = x80_from_x80p(S);
    }
}

// ---- the same accumulation, organised for throughput (what launch_pool_accum uses when it is given scratch):
//   * member cells of every group as an index list (pool_cells_kernel: ballot compaction, cell order kept),
//     so a thread visits only its group's cells instead of testing all N;
//   * x transposed once per call into a CELL-MAJOR scratch xc[c * G + g]: the 32 lanes of a warp (32
//     consecutive genes) read one 256-byte line per cell, no shared-memory staging, no block barriers;
//   * one WARP per (tile of 32 genes, group): 9 375 independent warps at K = 15, G = 20 000 instead of 625
//     blocks whose duration is set by their largest group and whose last 33 form a second wave;
//   * the loads do not depend on the 80-bit recursion: PA_U cells are in flight ahead of the additions.
// Same operations in the same (cell) order on every accumulator: bit-identical states.
#define PA_U 8
__global__ void __launch_bounds__(32) pool_cells_kernel(const uint8_t *member, int64_t N, int32_t *cells, int32_t *count)
{
    const int k = blockIdx.x, lane = threadIdx.x;
    const uint8_t *mem = member + (int64_t)k * N;
    int32_t *out = cells + (int64_t)k * N;
    int32_t n = 0;
    for (int64_t c0 = 0; c0 < N; c0 += 32) {
        const int64_t c = c0 + lane;
        const bool in = c < N && mem[c] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) out[n + __popc(bal & ((1u << lane) - 1u))] = (int32_t)c;
        n += __popc(bal);
    }
    if (lane == 0) count[k] = n;
}

template <bool CENTRED>
__global__ void __launch_bounds__(128)
    pool_accum_cm_kernel(const double *xc, int64_t G, int64_t N, const int32_t *cells, const int32_t *count,
                         const int32_t *order, int32_t K, const x80 *mean80, x80 *state)
{
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int32_t k = order ? order[blockIdx.y] : (int32_t)blockIdx.y;
    const int64_t g = tile * 32 + lane;
    if (tile * 32 >= G) return;
    const bool live = g < G;
    const int64_t gi = live ? g : G - 1; // lanes past the last gene repeat it (uniform loop, result dropped)
    const int64_t si = (int64_t)k * G + gi;
    constexpr bool centred = CENTRED;
    x80p S = x80p_from_x80(state[si]);
    const x80p M = x80p_from_x80(centred ? x80_neg(mean80[si]) : x80_zero());
    const int32_t *cl = cells + (int64_t)k * N;
    const int32_t n = count[k];
    const double *col = xc + gi;
    double v[PA_U], w[PA_U];
    int32_t i = 0;
#pragma unroll
    for (int u = 0; u < PA_U; u++) v[u] = u < n ? __ldcs(col + (int64_t)cl[u] * G) : 0.0;
    for (; i + PA_U <= n; i += PA_U) {
#pragma unroll
        for (int u = 0; u < PA_U; u++) {
            const int32_t j = i + PA_U + u;
            w[u] = j < n ? __ldcs(col + (int64_t)cl[j] * G) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PA_U; u++) {
            S = pool_step(S, M, v[u], centred);
            v[u] = w[u];
        }
    }
    for (int u = 0; i + u < n; u++) S = pool_step(S, M, v[u], centred);
    if (live) state[si] = x80_from_x80p(S);
}

// pass 0 closed: mean80 = state / n[k]
__global__ void pool_div_kernel(const x80 *state, const int32_t *n, int64_t G, int32_t K, x80 *mean80)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * G) return;
    mean80[i] = x80_div_u32(state[i], (uint32_t)n[i / G]);
}

// pass 1 closed: out = (double)(mean80 + state / n[k])
__global__ void pool_finish_kernel(const x80 *mean80, const x80 *state, const int32_t *n, int64_t G,
                                   int32_t K, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * G) return;
    out[i] = x80_to_double(x80_add(mean80[i], x80_div_u32(state[i], (uint32_t)n[i / G])));
}

cudaError_t launch_pool_accum(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                              int32_t K, const void *mean80, void *state, cudaStream_t st)
{
    pool_accum_kernel<<<(unsigned)((G + 31) / 32), dim3(32, PM_KB), 0, st>>>(
        x, ld, G, N, member, K, (const x80 *)mean80, (x80 *)state);
    return cudaGetLastError();
}

// bytes of scratch launch_pool_accum_fast needs: cell-major copy of x + cell lists + counts
size_t pool_accum_scratch_bytes(int64_t G, int64_t N, int32_t K)
{
    return (size_t)G * N * sizeof(double) + (size_t)K * N * sizeof(int32_t) + (size_t)K * sizeof(int32_t) + 512;
}

// prepare: transposes x and builds the cell lists (once per matrix / membership; both passes reuse them)
cudaError_t launch_pool_accum_prepare(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                                      int32_t K, void *scratch, cudaStream_t st)
{
    double *xc = (double *)scratch;
    int32_t *cells = (int32_t *)((char *)scratch + (((size_t)G * N * sizeof(double) + 255) & ~(size_t)255));
    int32_t *count = cells + (size_t)K * N;
    cudaError_t e = launch_transpose_f64_back(x, ld, G, N, xc, st);
    if (e != cudaSuccess) return e;
    pool_cells_kernel<<<(unsigned)K, 32, 0, st>>>(member, N, cells, count);
    return cudaGetLastError();
}

cudaError_t launch_pool_accum_fast(int64_t G, int64_t N, int32_t K, const int32_t *order, const void *mean80,
                                   void *state, const void *scratch, cudaStream_t st)
{
    const double *xc = (const double *)scratch;
    const int32_t *cells = (const int32_t *)((const char *)scratch + (((size_t)G * N * sizeof(double) + 255) & ~(size_t)255));
    const int32_t *count = cells + (size_t)K * N;
    const dim3 grid((unsigned)((G + 127) / 128), (unsigned)K);
    if (mean80)
        pool_accum_cm_kernel<true><<<grid, 128, 0, st>>>(xc, G, N, cells, count, order, K, (const x80 *)mean80, (x80 *)state);
    else
        pool_accum_cm_kernel<false><<<grid, 128, 0, st>>>(xc, G, N, cells, count, order, K, nullptr, (x80 *)state);
    return cudaGetLastError();
}
cudaError_t launch_pool_div(const void *state, const int32_t *n, int64_t G, int32_t K, void *mean80,
                            cudaStream_t st)
{
    const int64_t tot = (int64_t)K * G;
    pool_div_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const x80 *)state, n, G, K, (x80 *)mean80);
    return cudaGetLastError();
}
cudaError_t launch_pool_finish(const void *mean80, const void *state, const int32_t *n, int64_t G,
                               int32_t K, double *out, cudaStream_t st)
{
    const int64_t tot = (int64_t)K * G;
    pool_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const x80 *)mean80, (const x80 *)state,
                                                                      n, G, K, out);
    return cudaGetLastError();
}

// The whole mean on one GPU: `scratch` holds 2 * K * G 16-byte records; `fast` (pool_accum_scratch_bytes, or
// NULL) the cell-major copy and cell lists of the throughput form; `order` (device [K] or NULL): groups by
// decreasing size, so that the longest accumulations start first.
cudaError_t launch_pool_mean(const double *x, int64_t ld, int64_t G, int64_t N,
                             const uint8_t *member, const int32_t *gsize, int32_t K, double *out,
                             void *scratch, void *fast, const int32_t *order, cudaStream_t st)
{
    const size_t rec = (size_t)K * G * sizeof(x80);
    void *state = scratch, *mean80 = (char *)scratch + rec;
    cudaError_t e = cudaMemsetAsync(state, 0, rec, st);
    if (e != cudaSuccess) return e;
    if (fast && (e = launch_pool_accum_prepare(x, ld, G, N, member, K, fast, st)) != cudaSuccess) return e;
    auto accum = [&](const void *m80) {
        return fast ? launch_pool_accum_fast(G, N, K, order, m80, state, fast, st)
                    : launch_pool_accum(x, ld, G, N, member, K, m80, state, st);
    };
    if ((e = accum(nullptr)) != cudaSuccess) return e;
    if ((e = launch_pool_div(state, gsize, G, K, mean80, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(state, 0, rec, st)) != cudaSuccess) return e;
    if ((e = accum(mean80)) != cudaSuccess) return e;
    return launch_pool_finish(mean80, state, gsize, G, K, out, st);
}

// sd(x) of each pooled vector (cov.c cov_complete1 replayed): one warp per vector; the lanes form the
// per-element terms in parallel and add each block of 32 to the running 80-bit sum in order
// (x80_block_add_warp).
__global__ void __launch_bounds__(32) pooled_sd_kernel(const double *pooled, int64_t G, double *sd)
{
    const double *xv = pooled + (int64_t)blockIdx.x * G;
    const int lane = threadIdx.x;
    x80 acc = x80_zero(), mean80 = x80_zero(), xm = x80_zero();
    for (int pass = 0; pass < 3; pass++) {
        acc = x80_zero();
        // the term of element i in this pass (does not depend on the running sum: the terms of block i + 1 are formed
        // while block i is being added, which hides their latency behind the scan's dependent chain)
        auto make = [&](int64_t i) {
            x80 term = x80_zero();
            if (i < G) {
                x80 v = x80_from_double(xv[i]);
                if (pass == 0)
                    term = v;
                else if (pass == 1)
                    term = x80_sub(v, mean80);
                else {
                    x80 d = x80_sub(v, xm);
                    term = x80_mul(d, d);
                }
            }
            return term;
        };
        x80 next = make(lane);
        for (int64_t i0 = 0; i0 < G; i0 += 32) {
            const x80 term = next;
            next = make(i0 + 32 + lane);
            acc = x80_block_add_warp(acc, term, lane, (int)((G - i0) < 32 ? (G - i0) : 32));
        }
        if (pass == 0) {
            mean80 = x80_div_u32(acc, (uint32_t)G);
        } else if (pass == 1) {
            mean80 = x80_add(mean80, x80_div_u32(acc, (uint32_t)G));
            xm = x80_from_double(x80_to_double(mean80)); // xm[i] = (double)tmp
        } else if (lane == 0) {
            double var = x80_to_double(x80_div_u32(acc, (uint32_t)(G - 1)));
            sd[blockIdx.x] = sqrt(var);
        }
    }
}

cudaError_t launch_pooled_sd(const double *pooled, int64_t G, int32_t K, double *sd, cudaStream_t st)
{
    pooled_sd_kernel<<<(unsigned)K, 32, 0, st>>>(pooled, G, sd);
    return cudaGetLastError();
}

// out[k*G+g] = #{c : member[k*N+c] && v[g*ld+c] > 0}.  One warp per SNP row, lanes stride the cells
// (coalesced 128 B rows of int32), groups as a bitmask per cell (K <= 32 per launch).  Each group's
// count is a ballot + popc per 32 cells, so the kernel is a pure stream over the matrix:
// 4 B / cell read once, K*4 B / row written.  Integer, exact.
#define PC_WARPS 8
__global__ void __launch_bounds__(32 * PC_WARPS)
    pool_count_kernel(const int32_t *v, int64_t ld, int64_t G, int64_t N, const uint32_t *cellmask,
                      int32_t K, int32_t *out)
{
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * PC_WARPS + (threadIdx.x >> 5);
    if (g >= G) return;
    const int32_t *row = v + g * ld;
    int32_t cnt = 0; // lane k accumulates group k
    for (int64_t c0 = 0; c0 < N; c0 += 32) {
        const int64_t c = c0 + lane;
        uint32_t m = 0u;
        if (c < N && ld_stream(row + c) > 0) m = cellmask[c];
        for (int k = 0; k < K; k++) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (m >> k) & 1u);
            if (lane == k) cnt += __popc(bal);
        }
    }
    if (lane < K) out[(int64_t)lane * G + g] = cnt;
}

// Same result, 16-byte loads: lane l takes cells 128 s + 4 l .. + 3 of slab s, four ballots give the
// positive cells of the slab as four 32-bit words, and lane k counts group k with one AND + POPC per word
// against the group's membership words in the same bit order (built once per block in shared memory).
// ~6 instructions per 128 bytes instead of ~35: the kernel becomes a plain HBM stream.
__global__ void __launch_bounds__(256)
    pool_count_v4_kernel(const int32_t *v, int64_t ld, int64_t G, int64_t N, const uint32_t *cellmask,
                         int32_t K, int32_t *out)
{
    extern __shared__ uint4 pc_w[]; // [nslab][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t nslab = (N + 127) / 128;
    for (int64_t s = warp; s < nslab; s += nw) {
        uint32_t m[4], w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t c = 128 * s + 4 * lane + i;
            m[i] = c < N ? cellmask[c] : 0u;
        }
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t bal = __ballot_sync(0xffffffffu, (m[i] >> k) & 1u);
                if (lane == k) w[i] = bal;
            }
        if (lane < K) pc_w[s * K + lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    for (int64_t g = (int64_t)blockIdx.x * nw + warp; g < G; g += (int64_t)gridDim.x * nw) {
        const int32_t *row = v + g * ld;
        int32_t cnt = 0;
#pragma unroll 2
        for (int64_t s = 0; s < nslab; s++) {
            const int64_t c = 128 * s + 4 * lane;
            int4 x = make_int4(0, 0, 0, 0);
            // bounded by N, not ld: for a column-slice view (ld = the parent's stride) the tail of the
            // last row beyond N lies outside the allocation
            if (c + 3 < N) {
                x = __ldcs(reinterpret_cast<const int4 *>(row + c));
            } else {
                if (c < N) x.x = row[c];
                if (c + 1 < N) x.y = row[c + 1];
                if (c + 2 < N) x.z = row[c + 2];
            }
            const uint32_t b0 = __ballot_sync(0xffffffffu, x.x > 0), b1 = __ballot_sync(0xffffffffu, x.y > 0),
                           b2 = __ballot_sync(0xffffffffu, x.z > 0), b3 = __ballot_sync(0xffffffffu, x.w > 0);
            if (lane < K) {
                const uint4 w = pc_w[s * K + lane];
                cnt += __popc(b0 & w.x) + __popc(b1 & w.y) + __popc(b2 & w.z) + __popc(b3 & w.w);
            }
        }
        if (lane < K) out[(int64_t)lane * G + g] = cnt;
    }
}

cudaError_t launch_pool_count(const int32_t *v, int64_t ld, int64_t G, int64_t N,
                              const uint32_t *cellmask, int32_t K, int32_t *out, cudaStream_t st)
{
    const size_t smem = (size_t)((N + 127) / 128) * K * sizeof(uint4);
    if (((uintptr_t)v & 15) == 0 && ld % 4 == 0 && smem <= 200 * 1024) {
        if (smem > 48 * 1024) // per device: set whenever it is needed
            cudaFuncSetAttribute(pool_count_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = smem > 48 * 1024 ? 1 : (smem > 24 * 1024 ? 4 : 8);
        const unsigned grid = (unsigned)std::min<int64_t>((G + 7) / 8, (int64_t)sms * per_sm);
        pool_count_v4_kernel<<<grid, 256, smem, st>>>(v, ld, G, N, cellmask, K, out);
        return cudaGetLastError();
    }
    pool_count_kernel<<<(unsigned)((G + PC_WARPS - 1) / PC_WARPS), 32 * PC_WARPS, 0, st>>>(v, ld, G, N,
                                                                                          cellmask, K, out);
    return cudaGetLastError();
}

// Sub-matrix of a resident gene-major matrix: out[i*ldo + j] = x[rows[i]*ld + cols[j]] (NULL = identity).
// What a recursion level of the split-and-retest driver needs instead of `gexp.norm[rows, g1]` on the host
// (R/HoneyBADGER.R:1304-1309): index sets go down, the matrix stays on the device.
template <typename T>
__global__ void gather_rc_kernel(const T *x, int64_t ld, const int32_t *rows, int64_t nrows, const int32_t *cols,
                                 int64_t ncols, T *out, int64_t ldo)
{
    const int64_t i = blockIdx.x;
    const int64_t r = rows ? rows[i] : i;
    const T *src = x + r * ld;
    T *dst = out + i * ldo;
    for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < ncols; j += (int64_t)gridDim.y * blockDim.x)
        dst[j] = src[cols ? cols[j] : j];
}

cudaError_t launch_gather_rc(const void *x, int32_t elem, int64_t ld, const int32_t *rows, int64_t nrows,
                             const int32_t *cols, int64_t ncols, void *out, int64_t ldo, cudaStream_t st)
{
    const dim3 grid((unsigned)nrows, (unsigned)std::min<int64_t>((ncols + 255) / 256, 64));
    if (elem == 8)
        gather_rc_kernel<double><<<grid, 256, 0, st>>>((const double *)x, ld, rows, nrows, cols, ncols, (double *)out, ldo);
    else
        gather_rc_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t *)x, ld, rows, nrows, cols, ncols, (int32_t *)out, ldo);
    return cudaGetLastError();
}

// cellmask[c] = OR_k (member[k*N+c] != 0) << k
__global__ void build_cellmask_kernel(const uint8_t *member, int64_t N, int32_t K, uint32_t *mask)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    uint32_t m = 0;
    for (int k = 0; k < K; k++) m |= (member[(int64_t)k * N + c] ? 1u : 0u) << k;
    mask[c] = m;
}

cudaError_t launch_build_cellmask(const uint8_t *member, int64_t N, int32_t K, uint32_t *mask,
                                  cudaStream_t st)
{
    build_cellmask_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(member, N, K, mask);
    return cudaGetLastError();
}

// =====================================================================================
// Input construction (hb_inputs.cuh)
cudaError_t launch_row_means(const double *x, int64_t ld, int64_t G, int64_t N, double *out, cudaStream_t st)
{
    row_means_kernel<<<(unsigned)((G + 31) / 32), 128, 0, st>>>(x, ld, G, N, out);
    return cudaGetLastError();
}
cudaError_t launch_col_stats(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *keep,
                             double *center, double *scl, cudaStream_t st)
{
    col_stats_kernel<<<(unsigned)((N + 31) / 32), 32, 0, st>>>(x, ld, G, N, keep, center, scl);
    return cudaGetLastError();
}
cudaError_t launch_gexp_apply(const double *x, int64_t ld, int64_t N, const int32_t *rows, int64_t Gk,
                              const double *center, const double *scl, const double *refmean, double *out,
                              int64_t ldo, cudaStream_t st)
{
    const unsigned gy = (unsigned)std::min<int64_t>((N + 255) / 256, 64);
    gexp_apply_kernel<<<dim3((unsigned)Gk, gy), 256, 0, st>>>(x, ld, N, rows, Gk, center, scl, refmean, out, ldo);
    return cudaGetLastError();
}
cudaError_t launch_nz_sum(const double *x, int64_t ld, int64_t G, int64_t N, int nblocks, double *part_hi,
                          double *part_lo, unsigned long long *part_n, double *part_max, cudaStream_t st)
{
    nz_sum_kernel<<<nblocks, 256, 0, st>>>(x, ld, G, N, part_hi, part_lo, part_n, part_max);
    return cudaGetLastError();
}
cudaError_t launch_nz_mean_seq(const double *x, int64_t ld, int64_t G, int64_t N, unsigned long long n,
                               double *out, cudaStream_t st)
{
    nz_mean_seq_kernel<<<1, 32, 0, st>>>(x, ld, G, N, n, out);
    return cudaGetLastError();
}
cudaError_t launch_allele_flip(const int32_t *r, const int32_t *n, int64_t ld, int64_t N, const int32_t *rows,
                               int64_t Gk, const uint8_t *code, int32_t *out_r, int32_t *out_n, int64_t ldo,
                               cudaStream_t st)
{
    const unsigned gy = (unsigned)std::min<int64_t>((N + 255) / 256, 64);
    allele_flip_kernel<<<dim3((unsigned)Gk, gy), 256, 0, st>>>(r, n, ld, N, rows, code, out_r, out_n, ldo);
    return cudaGetLastError();
}

} // namespace hb
