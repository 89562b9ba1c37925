// hb_group.cuh — device grouping of cells: windowed smoothing, R dist(), Ward.D2 linkage.
//
// SURVEY.md 8f-2, the step right BEFORE the HMM path: it supplies the cut-tree group labels the
// pooled rounds run on.  Reference: R/HoneyBADGER.R:1122-1136 (expression: caTools::runmean k = 101,
// dist, NA -> 0, hclust "ward.D2", cutree) and :1378-1397 (allele: r.maf / n.sc, k = 31).  In R this
// is O(N^2 G) on one core and stops scaling near 10^4 cells; here it is three kernels.
// Group MEMBERSHIP is what the path consumes; distances are not compared bit for bit.
#pragma once
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

} // namespace hb
