// hb_group.cuh — device grouping of cells: windowed smoothing, R dist(), Ward.D2 linkage.
//
// SURVEY.md 8f-2, the step right BEFORE the HMM path: it supplies the cut-tree group labels the
// pooled rounds run on.  Reference: R/HoneyBADGER.R:1122-1136 (expression: caTools::runmean k = 101,
// dist, NA -> 0, hclust "ward.D2", cutree) and :1378-1397 (allele: r.maf / n.sc, k = 31).  In R this
// is O(N^2 G) on one core and stops scaling near 10^4 cells; here it is three kernels.
// Group MEMBERSHIP is what the path consumes; distances are not compared bit for bit.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace hb {

// ---------------------------------------------------------------------------------------------
// caTools::runmean(x, k) down the rows of every column: centred window of k (odd) positions that
// shrinks at both ends (endrule "mean"), non-finite values skipped, NaN where the window is empty.
// One thread per (position, cell); the window is summed afresh in row order (coalesced along cells,
// the overlap between neighbouring windows is served by L1/L2).  RATIO: the input is r / n of two
// int32 count planes (NaN where n == 0), as R/HoneyBADGER.R:1379.
template <bool RATIO>
__global__ void __launch_bounds__(256)
    runmean_kernel(const double *x, const int32_t *r, const int32_t *n, int64_t ld, int64_t G, int64_t N,
                   int32_t half, double *out, int64_t ldo)
{
    const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t g = blockIdx.x; // positions on grid.x (200 k SNPs exceed grid.y)
    if (c >= N) return;
    const int64_t lo = g - half < 0 ? 0 : g - half, hi = g + half >= G ? G - 1 : g + half;
    double s = 0.0;
    int32_t cnt = 0;
    for (int64_t t = lo; t <= hi; t++) {
        double v;
        if (RATIO) {
            const int32_t nn = n[t * ld + c];
            v = nn > 0 ? (double)r[t * ld + c] / (double)nn : NAN;
        } else {
            v = x[t * ld + c];
        }
        if (isfinite(v)) {
            s += v;
            cnt++;
        }
    }
    out[g * ldo + c] = cnt ? s / (double)cnt : NAN;
}

// r / n of two int32 count planes as doubles (NaN where n == 0), once per element: runmean_kernel<true> forms the
// ratio again for each of the k positions of every window — k fp64 divisions per output; with the ratios in a
// scratch matrix the windows are plain sums (same divisions, same order of additions: identical results).
__global__ void __launch_bounds__(256)
    ratio_kernel(const int32_t *r, const int32_t *n, int64_t ld, int64_t G, int64_t N, double *out, int64_t ldo)
{
    const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t g = blockIdx.x;
    if (c >= N) return;
    const int32_t nn = n[g * ld + c];
    out[g * ldo + c] = nn > 0 ? (double)r[g * ld + c] / (double)nn : NAN;
}

// ---------------------------------------------------------------------------------------------
// stats::dist(t(S)) squared, Euclidean, with R's NA rule: pairs with a non-finite member are left
// out and the sum is scaled up by G / (pairs used); no usable pair -> NA, which the reference then
// sets to 0 (R/HoneyBADGER.R:1126-1128).  Direct differences (as R's C code), 64 x 64 tile of cell
// pairs per block, 4 x 4 per thread (rows ty + 16 a, columns tx + 16 b: the 16 lanes of a half warp read 16
// consecutive doubles, no bank conflicts), positions streamed through shared memory 16 at a time.
#define GD_T 64
#define GD_K 16
__global__ void __launch_bounds__(256)
    dist_sq_kernel(const double *S, int64_t ld, int64_t G, int64_t N, double *D)
{
    __shared__ double sa[GD_K][GD_T + 1], sb[GD_K][GD_T + 1];
    const int64_t i0 = (int64_t)blockIdx.y * GD_T, j0 = (int64_t)blockIdx.x * GD_T;
    if (j0 + GD_T <= i0) return; // lower triangle: mirrored from the upper one
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4], cnt[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = cnt[a][b] = 0.0;
    double nfast = 0.0; // positions of slabs without a single non-finite value: counted once, not per pair
    for (int64_t g0 = 0; g0 < G; g0 += GD_K) {
        int bad = 0;
        for (int e = threadIdx.x; e < GD_K * GD_T; e += 256) {
            const int k = e / GD_T, cc = e % GD_T;
            const int64_t g = g0 + k;
            const double a = (g < G && i0 + cc < N) ? S[g * ld + i0 + cc] : NAN;
            const double b = (g < G && j0 + cc < N) ? S[g * ld + j0 + cc] : NAN;
            bad |= !isfinite(a) || !isfinite(b);
            sa[k][cc] = a;
            sb[k][cc] = b;
        }
        // R_euclidean (distance.c): dev = x_i - x_j; if (!ISNAN(dev)) { dist += dev * dev; count++; } —
        // a product and an addition, each rounded (no FMA), positions in order
        if (!__syncthreads_or(bad)) {
#pragma unroll 4
            for (int k = 0; k < GD_K; k++) {
                double va[4], vb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    va[a] = sa[k][ty + 16 * a];
                    vb[a] = sb[k][tx + 16 * a];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const double d = va[a] - vb[b];
                        acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(d, d));
                    }
            }
            nfast += (double)GD_K;
        } else {
#pragma unroll 4
            for (int k = 0; k < GD_K; k++) {
                double va[4], vb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    va[a] = sa[k][ty + 16 * a];
                    vb[a] = sb[k][tx + 16 * a];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const double d = va[a] - vb[b]; // NaN or Inf-Inf when a member is unusable
                        if (isfinite(d)) {
                            acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(d, d));
                            cnt[a][b] += 1.0;
                        }
                    }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
            if (i < N && j < N) {
                // if (count != nc) dist /= ((double)count / nc); return sqrt(dist); hclust squares it again
                const double c = cnt[a][b] + nfast;
                double v = 0.0;
                if (c > 0) {
                    v = c != (double)G ? acc[a][b] / (c / (double)G) : acc[a][b];
                    const double r = sqrt(v);
                    v = __dmul_rn(r, r);
                }
                const double w = (i == j) ? 0.0 : v;
                D[i * N + j] = w;
                D[j * N + i] = w;
            }
        }
}

// ---------------------------------------------------------------------------------------------
// The same distances with the loads and the missing-value bookkeeping taken off the critical path.
// dist_sq_kernel stages a slab of 16 positions through shared memory with ordinary loads between two
// block barriers: while a block waits for its slab nothing of it computes; the test for non-finite
// values re-reads every staged element; and in a slab that holds one, every (pair, position) pays
// an isfinite test and a second fp64 addition for the pair count — 2.3x the cost of a clean slab,
// which is the normal case of the allele matrices (r / n is NaN wherever a cell has no coverage).
// Here
//   * a pre-pass (dist_mask_kernel) writes, once per matrix, a 16-bit mask of the finite positions of
//     every (slab of 16 positions, cell) and a byte per (slab, tile of 64 cells) that says whether
//     anything is missing at all;
//   * the pair count of a slab is popc(mask_i & mask_j) — 16 integer instructions per thread and slab
//     instead of 256 fp64 additions — and the sum takes a position iff its bit is set in that AND (a
//     difference whose bit is clear is turned into a denormal, whose square is +0), so a slab with
//     missing values costs the same three fp64 instructions per (pair, position) as a clean one plus a
//     bit test and one 32-bit select;
//   * slabs move with 16-byte cp.async into a three-stage ring, two slabs ahead of the arithmetic,
//     one block barrier per slab;
//   * columns past N are clamped (their pairs are never stored), so edge tiles stay on the fast path.
// Per pair the same operations in the same order as dist_sq_kernel — results are bit-identical
// (tests/test_gpu_group.py).  Needs ld even and S 16-byte aligned (the launcher checks).
#define GP_STAGES 3
__global__ void __launch_bounds__(GD_T)
    dist_mask_kernel(const double *S, int64_t ld, int64_t G, int64_t N, uint16_t *masks, uint8_t *flags, int ntile)
{
    const int64_t g0 = (int64_t)blockIdx.x * GD_K, c = (int64_t)blockIdx.y * GD_T + threadIdx.x;
    uint32_t m = 0;
    int bad = g0 + GD_K > G; // a partial last slab takes the masked path: its rows past G have no bit
    if (c < N) {
#pragma unroll
        for (int k = 0; k < GD_K; k++)
            if (g0 + k < G) {
                const bool ok = isfinite(S[(g0 + k) * ld + c]);
                m |= ok ? 1u << k : 0u;
                bad |= !ok;
            }
    }
    masks[((int64_t)blockIdx.x * ntile + blockIdx.y) * GD_T + threadIdx.x] = (uint16_t)m;
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flags[(int64_t)blockIdx.x * ntile + blockIdx.y] = bad ? 1 : 0;
}

// SYM: one matrix, D is [N][N], only the upper-triangle tiles are computed and mirrored.  !SYM: the rows of D are
// the cells of A, its columns the cells of B (a row block of the full matrix when the cells are sharded over
// ranks: A = this rank's cells, B = another rank's), every tile computed, D[i * ldd + j].
template <bool SYM>
__global__ void __launch_bounds__(256)
    dist_sq_pipe_kernel(const double *Sa, int64_t lda, int64_t Na, const uint16_t *maska, const uint8_t *flagsa,
                        int ntile_a, const double *Sb, int64_t ldb, int64_t Nb, const uint16_t *maskb,
                        const uint8_t *flagsb, int ntile_b, int64_t G, double *D, int64_t ldd)
{
    extern __shared__ __align__(16) double gp_sm[]; // [GP_STAGES][2][GD_K][GD_T]
    const int64_t i0 = (int64_t)blockIdx.y * GD_T, j0 = (int64_t)blockIdx.x * GD_T;
    if (SYM && j0 + GD_T <= i0) return; // lower triangle: mirrored from the upper one
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t nslab = (G + GD_K - 1) / GD_K;
    const int64_t clamp_a = (Na - 1) & ~(int64_t)1, clamp_b = (Nb - 1) & ~(int64_t)1;
    auto issue = [&](int64_t s) {
        if (s < nslab) {
            double *base = gp_sm + (s % GP_STAGES) * (2 * GD_K * GD_T);
            for (int e = threadIdx.x; e < 2 * GD_K * GD_T / 2; e += 256) {
                const int which = e >> 9, k = (e >> 5) & (GD_K - 1), ch = e & 31;
                int64_t g = s * GD_K + k, c = (which ? j0 : i0) + 2 * ch;
                const int64_t cl = which ? clamp_b : clamp_a;
                g = g < G ? g : G - 1;
                c = c < cl ? c : cl;
                const double *src = which ? Sb + g * ldb + c : Sa + g * lda + c;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(base + which * (GD_K * GD_T) + k * GD_T + 2 * ch);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[4][4];
    int32_t cnt[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            acc[a][b] = 0.0;
            cnt[a][b] = 0;
        }
    int32_t nfull = 0; // positions of slabs without a missing value: counted once, not per pair
    const uint16_t *mrow = maska + (int64_t)blockIdx.y * GD_T + ty, *mcol = maskb + (int64_t)blockIdx.x * GD_T + tx;
    const int64_t mstride_a = (int64_t)ntile_a * GD_T, mstride_b = (int64_t)ntile_b * GD_T;
    issue(0);
    issue(1);
    for (int64_t s = 0; s < nslab; s++) {
        const bool bad = flagsa[s * ntile_a + blockIdx.y] | flagsb[s * ntile_b + blockIdx.x];
        uint32_t mi[4], mj[4];
        if (bad) {
#pragma unroll
            for (int a = 0; a < 4; a++) {
                mi[a] = mrow[s * mstride_a + 16 * a];
                mj[a] = mcol[s * mstride_b + 16 * a];
            }
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads(); // slab s has landed for everybody, and everybody is done with slab s - 1
        issue(s + 2);    // into the stage slab s - 1 used
        const double *sa = gp_sm + (s % GP_STAGES) * (2 * GD_K * GD_T), *sb = sa + GD_K * GD_T;
        if (!bad) {
            // R_euclidean (distance.c): dev = x_i - x_j; dist += dev * dev — a product and an addition, each
            // rounded (no FMA), positions in order
#pragma unroll 4
            for (int k = 0; k < GD_K; k++) {
                double va[4], vb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    va[a] = sa[k * GD_T + ty + 16 * a];
                    vb[a] = sb[k * GD_T + tx + 16 * a];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const double d = va[a] - vb[b];
                        acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(d, d));
                    }
            }
            nfull += GD_K;
        } else {
            uint32_t m[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    m[a][b] = mi[a] & mj[b];
                    cnt[a][b] += __popc(m[a][b]);
                }
#pragma unroll
            for (int k = 0; k < GD_K; k++) {
                double va[4], vb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    va[a] = sa[k * GD_T + ty + 16 * a];
                    vb[a] = sb[k * GD_T + tx + 16 * a];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        // A missing member makes the difference NaN (or +-Inf).  Clearing the HIGH word of an
                        // unwanted difference leaves a denormal whose square rounds to +0, and acc + 0 == acc
                        // bit for bit (acc is a sum of squares, never -0): one integer select per (pair,
                        // position) instead of a 64-bit select after the add.
                        const double d = va[a] - vb[b];
                        int hi; // and-with-immediate into a predicate, one select (the C form becomes shift, shift, and)
                        asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
                            "selp.b32 %0, %3, 0, p;\n\t}"
                            : "=r"(hi)
                            : "r"(m[a][b]), "r"(1u << k), "r"(__double2hiint(d)));
                        const double dm = __hiloint2double(hi, __double2loint(d));
                        acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(dm, dm));
                    }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
            if (i < Na && j < Nb) {
                // if (count != nc) dist /= ((double)count / nc); return sqrt(dist); hclust squares it again
                const double c = (double)(cnt[a][b] + nfull);
                double v = 0.0;
                if (c > 0) {
                    v = c != (double)G ? acc[a][b] / (c / (double)G) : acc[a][b];
                    const double r = sqrt(v);
                    v = __dmul_rn(r, r);
                }
                if (SYM) {
                    const double w = (i == j) ? 0.0 : v;
                    D[i * ldd + j] = w;
                    D[j * ldd + i] = w;
                } else {
                    D[i * ldd + j] = v; // a cell against itself comes out as 0 by the arithmetic
                }
            }
        }
}

// ---------------------------------------------------------------------------------------------
// hclust(d, "ward.D2"): nearest-neighbour chain with the Lance-Williams update on SQUARED distances,
// one persistent block (the merges are sequential; every step is a row-wide argmin or update that
// 1024 threads do cooperatively; no host round trips).  Merge m joins the clusters that contain
// leaves ma[m] and mb[m] at height sqrt(d2); merges come out in chain order (children always before
// their parents), the host sorts them by height to cut the tree.  D is destroyed.
#define WL_THREADS 1024
// (value, index) ordering of the nearest-neighbour search: smaller distance first; on equal distance
// the previous chain element wins (this is what makes the chain terminate), then the lowest index.
__device__ __forceinline__ bool wl_better(double v1, int32_t k1, double v2, int32_t k2, int32_t prefer)
{
    if (k1 < 0) return false;
    if (k2 < 0) return true;
    if (v1 != v2) return v1 < v2;
    if (k1 == prefer) return true;
    if (k2 == prefer) return false;
    return k1 < k2;
}

__device__ __forceinline__ void wl_argmin(const double *row, const uint8_t *active, int64_t N, int32_t self,
                                          int32_t prefer, double *sval, int32_t *sidx, int32_t &best)
{
    double bv = INFINITY;
    int32_t bi = -1;
    for (int64_t k = threadIdx.x; k < N; k += WL_THREADS) {
        if (active[k] && k != self) {
            const double v = row[k];
            if (wl_better(v, (int32_t)k, bv, bi, prefer)) {
                bv = v;
                bi = (int32_t)k;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        const int32_t oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (wl_better(ov, oi, bv, bi, prefer)) {
            bv = ov;
            bi = oi;
        }
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sval[w] = bv;
        sidx[w] = bi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        bv = sval[threadIdx.x];
        bi = sidx[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int32_t oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (wl_better(ov, oi, bv, bi, prefer)) {
                bv = ov;
                bi = oi;
            }
        }
        if (threadIdx.x == 0) sidx[0] = bi;
    }
    __syncthreads();
    best = sidx[0];
    __syncthreads();
}

__global__ void __launch_bounds__(WL_THREADS)
    ward_nnchain_kernel(double *D, int64_t N, int32_t *size, uint8_t *active, int32_t *chain, int32_t *ma,
                        int32_t *mb, double *height)
{
    __shared__ double sval[WL_THREADS / 32];
    __shared__ int32_t sidx[WL_THREADS / 32];
    for (int64_t k = threadIdx.x; k < N; k += WL_THREADS) {
        size[k] = 1;
        active[k] = 1;
    }
    __syncthreads();
    int32_t clen = 0, nmerge = 0, first_active = 0;
    while (nmerge < N - 1) {
        if (clen == 0) {
            while (!active[first_active]) first_active++;
            if (threadIdx.x == 0) chain[0] = first_active;
            clen = 1;
            __syncthreads();
        }
        for (;;) {
            const int32_t a = chain[clen - 1];
            const int32_t prev = clen >= 2 ? chain[clen - 2] : -1;
            int32_t b;
            wl_argmin(D + (int64_t)a * N, active, N, a, prev, sval, sidx, b);
            if (b == prev) break;
            if (threadIdx.x == 0) chain[clen] = b;
            clen++;
            __syncthreads();
        }
        const int32_t a = chain[clen - 1], b = chain[clen - 2]; // reciprocal nearest neighbours
        clen -= 2;
        const double dab = D[(int64_t)a * N + b];
        const int32_t na = size[a], nb = size[b];
        __syncthreads();
        // the union lives in slot b; slot a dies
        for (int64_t k = threadIdx.x; k < N; k += WL_THREADS) {
            if (active[k] && k != a && k != b) {
                const double nk = (double)size[k];
                const double v = ((na + nk) * D[(int64_t)a * N + k] + (nb + nk) * D[(int64_t)b * N + k] - nk * dab) /
                                 (na + nb + nk);
                D[(int64_t)b * N + k] = v;
                D[k * N + b] = v;
            }
        }
        if (threadIdx.x == 0) {
            ma[nmerge] = a;
            mb[nmerge] = b;
            height[nmerge] = sqrt(dab > 0 ? dab : 0.0);
            size[b] = na + nb;
            active[a] = 0;
        }
        nmerge++;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// The same linkage on a thread-block CLUSTER (up to 16 SMs).  The merges stay sequential, but every
// step of the single-block kernel is bounded by one SM's load latency and issue rate over a row of
// N distances; here each CTA owns a contiguous slice of the columns and
//   * a thread keeps its (at most WC_OWN) columns' liveness in a register bit mask, so a row scan is
//     one round of independent L2 loads (no dependent `active[k]` load in front of each);
//   * the per-CTA winners are exchanged through distributed shared memory — every CTA stores its
//     (value, index) into all peers' slot tables, one hardware cluster barrier, every CTA reduces the
//     same table — so all CTAs take the same decision without a broadcast step;
//   * control state (chain, cluster sizes) is replicated per CTA (size_rep / chain_rep: [R][N]),
//     only the distance matrix is shared, read with ld.global.cg after a fence + cluster barrier.
// Same comparison order (wl_better is a total order, so the winner does not depend on how the
// reduction is split), same Lance-Williams expression: merges, order and heights are bit-identical
// to ward_nnchain_kernel's (tests/test_gpu_group.py).  One barrier per chain extension, two per
// merge.
#define WC_THREADS 256
#define WC_WARPS (WC_THREADS / 32)
#define WC_OWN 16
#define WC_MAXR 16
// distance -> unsigned key with the order of the doubles (-0 counted as +0, as `v1 != v2` does; no NaN here)
__device__ __forceinline__ uint64_t wc_key(double v)
{
    const uint64_t u = (uint64_t)__double_as_longlong(v + 0.0);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
// index -> key of the tie rule: the preferred cluster first, then the lowest index
__device__ __forceinline__ uint32_t wc_ikey(int32_t k, int32_t prefer) { return k == prefer ? 0u : (uint32_t)k + 1u; }
// lexicographic minimum of (key, ik) over the warp, result in every lane
__device__ __forceinline__ void wc_warp_min(uint64_t &key, uint32_t &ik)
{
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const uint32_t mi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? ik : 0xffffffffu);
    key = ((uint64_t)mh << 32) | ml;
    ik = mi;
}
// OWN: columns per thread the instantiation is unrolled for (>= what the launch needs; a tight bound keeps the
// per-step instruction count — and with it the step's latency — down: 362 warp-instructions per step at OWN = 16
// for a launch that needs 3)
template <int OWN>
__global__ void __launch_bounds__(WC_THREADS)
    ward_cluster_kernel(double *D, int64_t N, int32_t *size_rep, int32_t *chain_rep, int32_t *ma, int32_t *mb,
                        double *height)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int R = (int)cl.num_blocks(), rk = (int)cl.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // winners of every warp of every CTA, two tables used alternately
    __shared__ uint64_t pkey[2][WC_MAXR * WC_WARPS];
    __shared__ uint32_t pidx[2][WC_MAXR * WC_WARPS];
    int32_t *size = size_rep + (int64_t)rk * N, *chain = chain_rep + (int64_t)rk * N;
    const int64_t chunk = ((N + R - 1) / R + 31) & ~(int64_t)31;
    const int64_t lo = (int64_t)rk * chunk, hi = lo + chunk < N ? lo + chunk : N;
    uint32_t alive = 0;
#pragma unroll
    for (int i = 0; i < OWN; i++)
        if (lo + tid + (int64_t)i * WC_THREADS < hi) alive |= 1u << i;
    for (int64_t k = tid; k < N; k += WC_THREADS) size[k] = 1;
    __syncthreads();
    cl.sync(); // every CTA of the cluster is running: its shared memory may be written from now on
    int par = 0;
    const int nent = R * WC_WARPS;

    const int nown = (int)((chunk + WC_THREADS - 1) / WC_THREADS); // columns per thread in use (uniform)

    // nearest active neighbour of cluster a (ties: `prefer`, then lowest index), known to every thread on return.
    // Candidates are compared as (wc_key(distance), wc_ikey(index)) pairs of unsigned integers — the same
    // total order as wl_better — so that a warp finds its winner with three redux.sync.min instead of five
    // rounds of 64-bit shuffles and fp64 compares.
    auto argmin = [&](int32_t a, int32_t prefer) -> int32_t {
        const double *row = D + (int64_t)a * N + lo + tid;
        double v[OWN];
#pragma unroll
        for (int i = 0; i < OWN; i++)
            if (i < nown && ((alive >> i) & 1u)) v[i] = __ldcg(row + (int64_t)i * WC_THREADS);
        uint64_t bk = ~0ull;
        uint32_t bi = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < OWN; i++) {
            if (i < nown) {
                const int32_t k = (int32_t)(lo + tid + (int64_t)i * WC_THREADS);
                if (((alive >> i) & 1u) && k != a) {
                    const uint64_t key = wc_key(v[i]);
                    const uint32_t ik = wc_ikey(k, prefer);
                    if (key < bk || (key == bk && ik < bi)) {
                        bk = key;
                        bi = ik;
                    }
                }
            }
        }
        wc_warp_min(bk, bi);
        if (lane < R) { // lane t hands this warp's winner to CTA t
            *cl.map_shared_rank(&pkey[par][rk * WC_WARPS + warp], lane) = bk;
            *cl.map_shared_rank(&pidx[par][rk * WC_WARPS + warp], lane) = bi;
        }
        cl.sync();
        bk = ~0ull;
        bi = 0xffffffffu;
        for (int e = lane; e < nent; e += 32) {
            const uint64_t key = pkey[par][e];
            const uint32_t ik = pidx[par][e];
            if (key < bk || (key == bk && ik < bi)) {
                bk = key;
                bi = ik;
            }
        }
        wc_warp_min(bk, bi);
        par ^= 1; // the next exchange uses the other table: a fast CTA cannot overwrite what a slow one still reads
        return bi == 0xffffffffu ? -1 : bi == 0 ? prefer : (int32_t)bi - 1;
    };

    int32_t clen = 0, nmerge = 0, first_active = 0;
    while (nmerge < N - 1) {
        if (clen == 0) {
            while (size[first_active] == 0) first_active++;
            if (tid == 0) chain[0] = first_active;
            clen = 1;
            __syncthreads();
        }
        int32_t a = chain[clen - 1], prev = clen >= 2 ? chain[clen - 2] : -1;
        for (;;) {
            const int32_t b = argmin(a, prev);
            if (b == prev) break;
            if (tid == 0) chain[clen] = b;
            clen++;
            prev = a;
            a = b;
        }
        const int32_t b = prev; // a and b are reciprocal nearest neighbours
        clen -= 2;
        const double dab = __ldcg(D + (int64_t)a * N + b);
        const int32_t na = size[a], nb = size[b];
        // the union lives in slot b; slot a dies
        {
            const double *ra = D + (int64_t)a * N + lo + tid, *rb = D + (int64_t)b * N + lo + tid;
            double va[OWN], vb[OWN];
#pragma unroll
            for (int i = 0; i < OWN; i++)
                if (i < nown && ((alive >> i) & 1u)) {
                    va[i] = __ldcg(ra + (int64_t)i * WC_THREADS);
                    vb[i] = __ldcg(rb + (int64_t)i * WC_THREADS);
                }
#pragma unroll
            for (int i = 0; i < OWN; i++) {
                const int64_t k = lo + tid + (int64_t)i * WC_THREADS;
                if (i < nown && ((alive >> i) & 1u) && k != a && k != b) {
                    const double nk = (double)size[k];
                    const double v = ((na + nk) * va[i] + (nb + nk) * vb[i] - nk * dab) / (na + nb + nk);
                    D[(int64_t)b * N + k] = v;
                    D[k * N + b] = v;
                }
                if (k == a) alive &= ~(1u << i);
            }
        }
        __syncthreads(); // size[a], size[b] have been read by everybody
        if (tid == 0) {
            if (rk == 0) {
                ma[nmerge] = a;
                mb[nmerge] = b;
                height[nmerge] = sqrt(dab > 0 ? dab : 0.0);
            }
            size[b] = na + nb;
            size[a] = 0;
        }
        nmerge++;
        cl.sync(); // release / acquire at cluster scope: the updated row and column of b are visible to their readers
    }
    cl.sync(); // nobody leaves while a peer may still store into its shared memory
}

} // namespace hb
