// hb_inputs.cuh — construction of the matrices the segmentation path consumes:
//   setGexpMats  (R/HoneyBADGER.R:124-188): rowMeans filter, scale(), gexp.sc - rowMeans(gexp.ref)
//   setAlleleMats (R/HoneyBADGER.R:486-614): in-silico bulk counts, heterozygosity filter, lesser-
//                 allele flip.
// All matrices are gene-major on the device (`x[g * ld + c]`), like everywhere else in the library.
//
// Arithmetic contract (base R, x86-64): rowMeans / colMeans keep one LDOUBLE accumulator per
// row / column, add the entries in index order and divide by the count (array.c do_colsum);
// scale.default centres with colMeans, then divides by sqrt(sum(v^2) / max(1, n-1)) where v^2 is a
// double vector and sum() accumulates it in LDOUBLE (summary.c rsum).  The 80-bit accumulations are
// replayed with hb_x87.cuh, every double operation is a separately rounded IEEE operation (no FMA
// contraction), so the normalised matrix is the one R builds, bit for bit.
#pragma once
#include "hb_device.cuh"
#include "hb_x87.cuh"
#include "hb_x87_pair.cuh"

namespace hb {

// ------------------------------------------------------------------------------------------------
// rowMeans(x): one thread per gene walks the cells in order.  The block's 32 x RM_TC tile is staged
// through shared memory with cp.async one tile ahead, so global reads are coalesced along the cell
// axis and overlap the (latency-bound) 80-bit additions.
#define RM_TC 64
__device__ __forceinline__ void cpa8_raw(double *dst_smem, const double *g)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g) : "memory");
}

__global__ void __launch_bounds__(128)
    row_means_kernel(const double *__restrict__ x, int64_t ld, int64_t G, int64_t N,
                     double *__restrict__ out)
{
    __shared__ double tile[2][32][RM_TC + 1];
    const int64_t g0 = (int64_t)blockIdx.x * 32;
    const int tid = threadIdx.x;
    auto issue = [&](int buf, int64_t c0) {
        for (int i = tid; i < 32 * RM_TC; i += 128) {
            const int r = i / RM_TC, cc = i % RM_TC;
            const int64_t g = g0 + r, c = c0 + cc;
            if (g < G && c < N) cpa8_raw(&tile[buf][r][cc], x + g * ld + c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    x80 S = x80_zero();
    issue(0, 0);
    int buf = 0;
    for (int64_t c0 = 0; c0 < N; c0 += RM_TC, buf ^= 1) {
        if (c0 + RM_TC < N) {
            issue(buf ^ 1, c0 + RM_TC);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (tid < 32 && g0 + tid < G) {
            const int lim = (int)((N - c0) < RM_TC ? (N - c0) : RM_TC);
            for (int cc = 0; cc < lim; cc++) S = x80_add(S, x80_from_double(tile[buf][tid][cc]));
        }
        __syncthreads();
    }
    if (tid < 32 && g0 + tid < G) out[g0 + tid] = x80_to_double(x80_div_u32(S, (uint32_t)N));
}

// ------------------------------------------------------------------------------------------------
// Column statistics of scale(): per cell, over the kept genes in order,
//   center = (double)(sum80(x) / n)                     colMeans(x, na.rm=TRUE)
//   scl    = sqrt((double)sum80((x - center)^2) / max(1, n-1))
// One thread per cell; x[g * ld + c] is coalesced across the warp's cells; loads run CS_U genes ahead.  10^4 cells
// are only 313 warps, so the kernel runs at the LATENCY of one 80-bit addition: the running sums are carried as
// pairs of doubles (hb_x87_pair.cuh; both passes add doubles), whose dependent chain is ~20 fp64 operations
// instead of the ~50 integer instructions and branches of x80_add.
#define CS_U 16
__global__ void __launch_bounds__(32)
    col_stats_kernel(const double *__restrict__ x, int64_t ld, int64_t G, int64_t N,
                     const uint8_t *__restrict__ keep, double *__restrict__ center,
                     double *__restrict__ scl)
{
    const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (c >= N) return;
    const double *col = x + c;
    x80p S = {0.0, 0.0};
    uint32_t n = 0;
    for (int64_t g0 = 0; g0 < G; g0 += CS_U) {
        double v[CS_U];
#pragma unroll
        for (int u = 0; u < CS_U; u++) v[u] = (g0 + u < G) ? ld_stream(col + (g0 + u) * ld) : 0.0;
#pragma unroll
        for (int u = 0; u < CS_U; u++)
            if (g0 + u < G && (!keep || keep[g0 + u])) {
                S = x80p_add_double(S, v[u]);
                n++;
            }
    }
    const double m = n ? x80_to_double(x80_div_u32(x80_from_x80p(S), n)) : 0.0;
    center[c] = m;
    if (!scl) return;
    S.h = S.l = 0.0;
    for (int64_t g0 = 0; g0 < G; g0 += CS_U) {
        double v[CS_U];
#pragma unroll
        for (int u = 0; u < CS_U; u++) v[u] = (g0 + u < G) ? ld_stream(col + (g0 + u) * ld) : 0.0;
#pragma unroll
        for (int u = 0; u < CS_U; u++)
            if (g0 + u < G && (!keep || keep[g0 + u])) {
                const double d = __dsub_rn(v[u], m);
                S = x80p_add_double(S, __dmul_rn(d, d));
            }
    }
    const double den = (double)(n > 1 ? n - 1 : 1);
    scl[c] = sqrt(__ddiv_rn(S.h, den)); // S.h is the running sum rounded to double
}

// out[k][c] = ((x[rows[k]][c] - center[c]) / scl[c]) - refmean[k]; each stage optional (null pointer).
__global__ void __launch_bounds__(256)
    gexp_apply_kernel(const double *__restrict__ x, int64_t ld, int64_t N,
                      const int32_t *__restrict__ rows, int64_t Gk, const double *__restrict__ center,
                      const double *__restrict__ scl, const double *__restrict__ refmean,
                      double *__restrict__ out, int64_t ldo)
{
    const int64_t k = blockIdx.x; // rows on grid.x (200 k SNP rows exceed grid.y)
    const int64_t g = rows ? rows[k] : k;
    const double rm = refmean ? refmean[k] : 0.0;
    for (int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x; c < N; c += (int64_t)gridDim.y * 256) {
        double v = ld_stream(x + g * ld + c);
        if (center) v = __dsub_rn(v, center[c]);
        if (scl) v = __ddiv_rn(v, scl[c]);
        if (refmean) v = __dsub_rn(v, rm);
        out[k * ldo + c] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// mean(x[x != 0]) (the default minMeanTest / minMeanRef, R/HoneyBADGER.R:124).  R's mean is a
// sequential two-pass LDOUBLE evaluation over the whole matrix — a serial chain of 2·n additions.
// Stage 1 (nz_sum_kernel): the exact-to-~2^-100 sum in double-double, in parallel, plus the count.
// The caller forms the mean from it and checks that no row mean lies within a safety margin of the
// threshold it is compared with; only then can the serial rounding pattern matter, and stage 2
// (nz_mean_seq_kernel: one warp replays R's loop in x87 arithmetic, a block of 32 terms at a time) is run instead.
__device__ __forceinline__ void dd_add(double &hi, double &lo, double v)
{
    const double s = __dadd_rn(hi, v);
    const double bb = __dsub_rn(s, hi);
    const double err = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(v, bb));
    hi = s;
    lo = __dadd_rn(lo, err);
}

__global__ void __launch_bounds__(256)
    nz_sum_kernel(const double *__restrict__ x, int64_t ld, int64_t G, int64_t N, double *part_hi,
                  double *part_lo, unsigned long long *part_n, double *part_max)
{
    __shared__ double sh[256], sl[256], sm[256];
    __shared__ unsigned long long sn[256];
    double hi = 0.0, lo = 0.0, mx = 0.0;
    unsigned long long n = 0;
    const int64_t tot = G * N;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < tot; i += (int64_t)gridDim.x * 256) {
        const int64_t g = i / N, c = i - g * N;
        const double v = ld_stream(x + g * ld + c);
        if (v != 0.0) {
            dd_add(hi, lo, v);
            n++;
            mx = fmax(mx, fabs(v));
        }
    }
    sh[threadIdx.x] = hi, sl[threadIdx.x] = lo, sn[threadIdx.x] = n, sm[threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double H = 0.0, L = 0.0, M = 0.0;
        unsigned long long C = 0;
        for (int i = 0; i < 256; i++) {
            dd_add(H, L, sh[i]);
            dd_add(H, L, sl[i]);
            C += sn[i];
            M = fmax(M, sm[i]);
        }
        part_hi[blockIdx.x] = H, part_lo[blockIdx.x] = L, part_n[blockIdx.x] = C, part_max[blockIdx.x] = M;
    }
}

// R's real_mean over the non-zero entries in column-major (cell, then gene) order; `n` is their count.
// One warp: 32 consecutive genes of a cell are one block of terms (zeros are exact zeros and change nothing),
// added to the 80-bit running sum in order as integer prefix sums while the sum stays in its binade
// (x80_block_add_warp, hb_x87.cuh); the next block's values are loaded before the current one is added.
__global__ void __launch_bounds__(32)
    nz_mean_seq_kernel(const double *__restrict__ x, int64_t ld, int64_t G, int64_t N,
                       unsigned long long n, double *out)
{
    const int lane = threadIdx.x;
    x80 S = x80_zero(), M = x80_zero();
    const int64_t nblk = (G + 31) / 32;
    for (int pass = 0; pass < 2; pass++) {
        double vnext = lane < G ? x[(int64_t)lane * ld] : 0.0;
        for (int64_t c = 0; c < N; c++)
            for (int64_t b = 0; b < nblk; b++) {
                const double v = vnext;
                // the block after this one (next genes of the cell, or the first genes of the next cell)
                int64_t nb = b + 1, nc = c;
                if (nb == nblk) {
                    nb = 0;
                    nc = c + 1;
                }
                const int64_t ng = nb * 32 + lane;
                vnext = (nc < N && ng < G) ? x[ng * ld + nc] : 0.0;
                x80 term = x80_zero();
                if (v != 0.0) {
                    term = x80_from_double(v);
                    if (pass) term = x80_sub(term, M);
                }
                S = x80_block_add_warp(S, term, lane, 32);
            }
        // n < 2^32 is checked by the caller
        if (pass == 0) {
            M = x80_div_u32(S, (uint32_t)n);
            S = x80_zero();
        } else if (lane == 0) {
            *out = x80_to_double(x80_add(M, x80_div_u32(S, (uint32_t)n)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// setAlleleMats: gather the kept SNP rows and apply the lesser-allele flip.
//   code[k] = 0: r.maf = r      (E <= 0.5)
//             1: r.maf = n - r  (E > 0.5)
//             2: r.maf = 0      (E is NA)
__global__ void __launch_bounds__(256)
    allele_flip_kernel(const int32_t *__restrict__ r, const int32_t *__restrict__ n, int64_t ld,
                       int64_t N, const int32_t *__restrict__ rows, const uint8_t *__restrict__ code,
                       int32_t *__restrict__ out_r, int32_t *__restrict__ out_n, int64_t ldo)
{
    const int64_t k = blockIdx.x;
    const int64_t g = rows ? rows[k] : k;
    const int cd = code[k];
    for (int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x; c < N; c += (int64_t)gridDim.y * 256) {
        const int32_t rv = ld_stream(r + g * ld + c), nv = ld_stream(n + g * ld + c);
        out_r[k * ldo + c] = cd == 0 ? rv : (cd == 1 ? nv - rv : 0);
        if (out_n) out_n[k * ldo + c] = nv;
    }
}

} // namespace hb
