/*
 * hb_b200.h — C ABI of libhb_b200.so: B200 (sm_100a) HMM segmentation for HoneyBADGER.
 *
 * The reference (an R package) has no FFI for this path: the seam is the R-level pair
 *     z <- HiddenMarkov::dthmm(x, Pi, delta, distn, pm[, pn, discrete=TRUE]);  y <- HiddenMarkov::Viterbi(z)
 * at four call sites (R/HoneyBADGER.R:1150-1151, :1409-1410, R/HoneyBADGER_gexp.R:503-504,
 * R/HoneyBADGER_allele.R:579-580), fed by the pooled statistics of R/HoneyBADGER.R:1141,1149
 * (apply(.,1,mean), sd) and :1404-1405 (rowSums(.>0)).  The entry points below are what a `.Call`
 * shim binds instead (see INTEGRATION.md and r/hb_shim.c).
 *
 * Conventions
 *   - plain C, no C++ exceptions, no torch types; every function returns HB_OK (0) or a negative
 *     code and hb_last_error() returns a thread-local message.
 *   - *_host functions take HOST pointers in R's native layout (column-major, one chain per
 *     column) and copy to/from the device internally.  The caller owns all buffers; inputs are
 *     never written.
 *   - *_dev functions take DEVICE pointers in the resident gene-major layout x[t*ld + b]
 *     (t = gene/SNP position, b = chain column) and enqueue on `stream` (a cudaStream_t cast to
 *     void*, NULL = legacy default stream).  The chain entry points (hb_hmm_*_dev, *_tiled_dev, the
 *     long-chain and retest *_dev calls, transposes / tiling, pooled counts) never wait for the stream:
 *     host tables (segment offsets, emission table) go up with cudaMemcpyAsync from a pinned arena and are
 *     cached per device.  Entry points that RETURN host results (hb_pack_counts_dev's overflow check,
 *     hb_pool_mean_dev's group sizes, hb_set_*_mats_dev, hb_mean_nonzero_dev, the pooled-round calls,
 *     *_stats_dev, hb_status_dev) wait for `stream` — each says so.
 *   - Streams: the library keeps ONE scratch workspace per device (back-pointers, checkpoints, tables).
 *     Calls may come from any stream; a call on a different stream than the previous one first makes its
 *     stream wait (cudaStreamWaitEvent) for the previous call's work, so two streams never overlap on the
 *     scratch — calls are serialised per device, not per stream.  The status word reports the LAST launch
 *     (hb_status_dev before the next launch if every launch's underflow flag matters).
 *   - Kernel-choice switches (hb_set_*_mode) are per calling thread.
 *   - A chain is one (column b, segment s) pair; segments are [seg_off[s], seg_off[s+1]) on the
 *     position axis (seg_off is a HOST array of nseg+1 offsets; nseg = 1 reproduces the
 *     reference's single concatenated chain, R/HoneyBADGER.R:1117-1121).
 *   - *_host posteriors: the S - 1 non-neutral planes cross PCIe; the neutral-state plane (state 2 of 3, state 2
 *     of 2) is rebuilt on the host as |1 - sum of the others| — the definition the speculate-and-verify /
 *     lean kernels use on the device — by the copy-out threads that also widen the states to R integers.
 *   - Viterbi states are 1-based like R's; state paths are bit-identical to the reference
 *     recursion (operation order of HiddenMarkov::Viterbi.dthmm, first index wins ties).
 *     Posteriors gamma = exp(logalpha + logbeta - LL) of HiddenMarkov::forwardback agree to 1e-9.
 *   - There is no CPU fallback: without a CUDA device every compute entry point returns
 *     HB_ERR_NODEVICE / HB_ERR_CUDA.
 */
#ifndef HB_B200_H
#define HB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_ARG (-1)       /* bad argument */
#define HB_ERR_CUDA (-2)      /* CUDA runtime / launch failure */
#define HB_ERR_UNDERFLOW (-3) /* HiddenMarkov::Viterbi's stop("Problems With Underflow") */
#define HB_ERR_NOMEM (-4)
#define HB_ERR_NODEVICE (-5)

#define HB_WANT_VITERBI 1 /* state path y */
#define HB_WANT_GAMMA 2   /* forward-backward posteriors */
#define HB_WANT_LL 4      /* log-likelihood per chain (alone: the forward sweep only) */

/* ------------------------------------------------------------------ library / device */
int hb_version(void);
const char *hb_last_error(void);
int hb_device_count(int *count);
int hb_set_device(int device);
/* Frees the per-device scratch cache (psi back-pointers, alpha checkpoints, staging). */
int hb_release_workspace(void);
/* *_dev entry points never synchronise; this waits for `stream` and returns HB_ERR_UNDERFLOW if a
 * chain of the last launch ended with nu == -Inf (HiddenMarkov::Viterbi's error), else HB_OK. */
int hb_status_dev(void *stream);
/* Host threads the pipelined hb_hmm_*_host calls use to copy results out of the pinned landing buffers, widen
 * the states to R integers and rebuild the neutral posterior plane.  0 (default): all but two of the machine's
 * cores, at most 16.  Several processes sharing one host (one per GPU) should divide the cores between them. */
int hb_set_host_threads(int32_t n);
/* Gaussian chains, kernel choice.  Two kernels compute the same results (identical Viterbi paths,
 * posteriors within 1e-11 of each other): the literal-arithmetic kernel, and the speculate-and-verify
 * kernel (eligible when the parameter block is the reference's symmetric one, delta[1] > 0 and no
 * chain is longer than 4096 positions).  mode 0 (default): speculative when eligible and the row
 * strides are multiples of 128 elements, else literal; 1: literal always; 3: speculative whenever
 * eligible; 2: speculative with EVERY chain re-run through its exact fallback (test hook). */
int hb_set_norm3_mode(int32_t mode);
/* Number of chains of the last Gaussian launch whose Viterbi recursion was re-run with the literal
 * reference arithmetic because a decision margin fell below the rounding bound (waits for `stream`). */
int hb_norm3_stats_dev(void *stream, int64_t *rerun_chains);
/* Which chains those were: up to `cap` entries seg * B + b (chain column b of segment seg) of the last
 * Gaussian launch, *n of them written (at most the first 16384 are recorded; waits for `stream`). */
int hb_norm3_flagged_dev(void *stream, int64_t *chains, int64_t cap, int64_t *n);
/* Which kernels the last chain launches on the current device went to (no synchronisation; host-side
 * bookkeeping): *norm3_fast = 1 speculate-and-verify kernel, 0 literal kernel, -1 none yet;
 * *binom2_lean = 1 lean plain-double kernel (general kernel behind its gate), 0 general kernel only.
 * Either pointer may be NULL.  Lets a caller (and the parity tests) assert which code path produced a result. */
int hb_last_kernels(int32_t *norm3_fast, int32_t *binom2_lean);
/* Bytes the last pipelined hb_hmm_*_host call moved host -> device and device -> host, and bytes of posterior
 * planes it rebuilt on the host instead of downloading (the neutral-state plane).  Any pointer may be NULL. */
int hb_last_host_io(int64_t *h2d_bytes, int64_t *d2h_bytes, int64_t *rebuilt_bytes);
/* Binomial chains, kernel choice.  0 (default): the lean plain-double kernel, with the general
 * relative-exponent kernel re-running the launch when some position carries evidence beyond 2^+-900
 * (both enqueued, no synchronisation); 1: general kernel only.  Same results either way. */
int hb_set_binom2_mode(int32_t mode);
/* 1 if the last binomial launch had to be redone by the general kernel (waits for `stream`). */
int hb_binom2_stats_dev(void *stream, int32_t *general_rerun);
/* Largest trial count served by the host-built binomial emission table (default 511, <= 4095). */
int hb_set_binom_lut_max(int32_t nmax);

/* ------------------------------------------------------------------ 3-state Gaussian chains
 * Replaces dthmm(x, Pi, delta, "norm", list(mean=c(pd,pn,pa), sd=c(sd,sd,sd))) -> Viterbi
 * (R/HoneyBADGER.R:1144-1151, R/HoneyBADGER_gexp.R:497-504) and adds forwardback posteriors.
 * mean[3]; one sd per chain shared by the 3 states: sd[b] (sd_per_seg=0) or sd[s*B+b] (=1);
 * Pi[9] row-major Pi[j*3+k] = P(j->k); delta[3].
 */
int hb_hmm_norm_host(const double *x /* [T x B] column-major */, int64_t T, int64_t B,
                     const int64_t *seg_off, int32_t nseg, const double *mean, const double *sd,
                     int32_t sd_per_seg, const double *Pi, const double *delta, int32_t flags,
                     int32_t *y /* [T x B] column-major, 1-based, or NULL */,
                     double *gamma /* [T x B x 3] column-major, or NULL */,
                     double *ll /* [B x nseg] column-major, or NULL */);

int hb_hmm_norm_dev(const double *x /* device, x[t*ld+b] */, int64_t ld, int64_t T, int64_t B,
                    const int64_t *seg_off /* host */, int32_t nseg, const double *mean,
                    const double *sd /* device */, int32_t sd_per_seg, const double *Pi,
                    const double *delta, int32_t flags, uint8_t *y /* device y[t*ldy+b] */,
                    int64_t ldy, double *gamma /* device gamma[(k*T+t)*ldg+b] */, int64_t ldg,
                    double *ll /* device ll[s*B+b] */, void *stream);

/* Same chains on the WARP-TILED resident layout: element (t, b) of a matrix lives at
 * ((b/32)*T + t)*32 + b%32 (tiles of 32 consecutive chains, each tile [T][32] contiguous), B is
 * padded to a multiple of 32 in every buffer; y and each gamma plane (plane stride T*pad32(B)) use
 * the same layout.  Every warp then streams contiguous memory (2 KB per 8-step chunk), which the
 * memory system serves ~12 % faster than 256-byte accesses at row stride.
 * hb_tile_*_dev / hb_untile_*_dev convert from / to the gene-major layout. 
 * The tiled buffers (x / counts, y) must be 16-byte aligned (cudaMalloc'd buffers are): tiles are staged in
 * 16-byte pieces. */
int hb_hmm_norm_tiled_dev(const double *x, int64_t T, int64_t B, const int64_t *seg_off /* host */,
                          int32_t nseg, const double *mean, const double *sd /* device */,
                          int32_t sd_per_seg, const double *Pi, const double *delta, int32_t flags,
                          uint8_t *y, double *gamma, double *ll /* device ll[s*B+b] */, void *stream);
int hb_hmm_binom_tiled_dev(const int32_t *x, const int32_t *size, int64_t T, int64_t B,
                           const int64_t *seg_off /* host */, int32_t nseg, const double *prob,
                           double rho, const double *Pi, const double *delta, int32_t flags,
                           uint8_t *y, double *gamma, double *ll, void *stream);
/* Posteriors only, CHAIN-RESIDENT and parallel along the gene axis (csrc/hb_norm3_resident.cuh): every chain
 * is held in shared memory by one block of 1-4 warps, x is read once and the posteriors written once (32 B
 * of HBM traffic per step against 46 for the two-sweep kernel).  Layout: CHAIN-MAJOR, x[b*T + t] — R's own
 * column-major [T x B], no transpose — and gamma[(k*B + b)*T + t] = R's [T x B x 3].  Symmetric parameter
 * block only (R/HoneyBADGER.R:1144-1150), segments of at most 6 200 positions; same posteriors as
 * hb_hmm_norm_dev(HB_WANT_GAMMA) within 1e-9.  A degenerate transfer matrix (emission ratios beyond the
 * double range inside one piece of a chain) raises the gate word hb_fbr_stats_dev reports; the caller then
 * falls back to hb_hmm_norm_dev.  No synchronisation, and none of the shared per-device scratch: the call may
 * run on a second stream beside a chain launch (e.g. the Viterbi-only kernel on the warp-tiled copy). */
int hb_forwardback_norm_cm_dev(const double *x, int64_t T, int64_t B, const int64_t *seg_off /* host */,
                               int32_t nseg, const double *mean, const double *sd /* device */,
                               int32_t sd_per_seg, const double *Pi, const double *delta, double *gamma,
                               void *stream);
/* *gate = 1 if the last chain-resident launch met a degenerate transfer matrix (waits for `stream`). */
int hb_fbr_stats_dev(void *stream, int32_t *gate);
/* Binomial chains on PACKED counts: one uint32 per position, (size << 16) | x, both < 65536
 * (hb_pack_counts_dev packs two int32 planes elementwise — any layout — and returns HB_ERR_ARG if
 * a count does not fit).  Halves the input bytes of the allele path: 4 B / position instead of 8. */
int hb_pack_counts_dev(const int32_t *x, const int32_t *size, int64_t count, uint32_t *out, void *stream);
int hb_hmm_binom_packed_tiled_dev(const uint32_t *xn, int64_t T, int64_t B, const int64_t *seg_off,
                                  int32_t nseg, const double *prob, double rho, const double *Pi,
                                  const double *delta, int32_t flags, uint8_t *y, double *gamma,
                                  double *ll, void *stream);
/* gene-major src[t*ld+b] -> tiled dst (elem = 1, 4 or 8 bytes), and back; device pointers */
int hb_tile_dev(const void *src, int64_t ld, int64_t T, int64_t B, int32_t elem, void *dst, void *stream);
int hb_untile_dev(const void *src, int64_t T, int64_t B, int32_t elem, void *dst, int64_t ld, void *stream);

/* ------------------------------------------------------------------ 2-state binomial chains
 * Replaces dthmm(mafl, Pi, delta, "binom", list(prob=c(pd,pn)), list(size=sizel), discrete=TRUE)
 * -> Viterbi (R/HoneyBADGER.R:1408-1410, R/HoneyBADGER_allele.R:578-580).  prob[2], Pi[4], delta[2].
 * rho >= 0 is the beta-binomial over-dispersion (extension, not in the reference); rho = 0 is the
 * reference's binomial.
 */
int hb_hmm_binom_host(const int32_t *x /* [T x B] column-major */, const int32_t *size,
                      int64_t T, int64_t B, const int64_t *seg_off, int32_t nseg,
                      const double *prob, double rho, const double *Pi, const double *delta,
                      int32_t flags, int32_t *y, double *gamma /* [T x B x 2] */, double *ll);

int hb_hmm_binom_dev(const int32_t *x /* device x[t*ld+b] */, const int32_t *size, int64_t ld,
                     int64_t T, int64_t B, const int64_t *seg_off /* host */, int32_t nseg,
                     const double *prob, double rho, const double *Pi, const double *delta,
                     int32_t flags, uint8_t *y, int64_t ldy, double *gamma, int64_t ldg,
                     double *ll, void *stream);

/* ------------------------------------------------------------------ layout + pooled statistics
 * R matrices are column-major [G x N] (one cell's genes contiguous); the resident layout is
 * gene-major [G][ld>=N].  hb_transpose_* convert on the device (both pointers device).
 */
int hb_transpose_f64_dev(const double *src_colmajor, int64_t G, int64_t N, double *dst,
                         int64_t ld, void *stream);
/* counts stored as R doubles -> int32 gene-major (values are rounded with rint). */
int hb_transpose_f64_to_i32_dev(const double *src_colmajor, int64_t G, int64_t N, int32_t *dst,
                                int64_t ld, void *stream);
int hb_transpose_u8_back_dev(const uint8_t *src_gene_major, int64_t ld, int64_t G, int64_t N,
                             int32_t *dst_colmajor, void *stream);
int hb_transpose_f64_back_dev(const double *src_gene_major, int64_t ld, int64_t G, int64_t N,
                              double *dst_colmajor, void *stream);

/* Column sd over positions per chain (R's sd(): n-1 denominator), device -> device sd[b] or
 * sd[s*B+b].  Double-double accumulation; rounds like R's long-double code to <= 1 ulp. */
int hb_chain_sd_dev(const double *x, int64_t ld, int64_t T, int64_t B, const int64_t *seg_off,
                    int32_t nseg, int32_t per_seg, double *sd, void *stream);

/* Pooled expression observation of R/HoneyBADGER.R:1141 for K groups at once:
 * out[k*G+g] = mean over cells c with member[k*N+c] != 0 of x[g*ld+c], R `mean` semantics
 * (x87 long-double two-pass accumulation in cell order, emulated exactly on the device). */
int hb_pool_mean_dev(const double *x, int64_t ld, int64_t G, int64_t N,
                     const uint8_t *member /* device [K][N] */, int32_t K,
                     double *out /* device [K][G] */, void *stream);
/* The same mean as a RELAY, for cells sharded over several GPUs.  R's mean is an order-dependent
 * 80-bit accumulation, so partial sums cannot be merged afterwards; instead an 80-bit state per
 * (group, gene) — `state`, K*G records of 16 bytes, device — is continued by each rank in turn
 * with ITS cells (cell order = rank order) and handed to the next rank:
 *   phase 0  state += x over this rank's members               (start from zeros on the first rank)
 *   phase 1  mean80 = state / n_total[k]                        (last rank; then broadcast mean80, zero state)
 *   phase 2  state += (x - mean80) over this rank's members
 *   phase 3  out[k*G+g] = (double)(mean80 + state / n_total[k]) (last rank; R's result bit for bit)
 * member is this rank's [K][N] slice (device), n_total (device int32[K]) the group sizes over all ranks.
 * hb_pool_mean_dev is phases 0-3 on one GPU. */
int hb_pool_mean_relay_dev(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                           int32_t K, const int32_t *n_total, int32_t phase, void *state,
                           void *mean80, double *out, void *stream);
/* sd(x) of each pooled vector, R semantics (long-double emulation): out_sd[k]. */
int hb_pooled_sd_dev(const double *pooled /* device [K][G] */, int64_t G, int32_t K,
                     double *out_sd /* device [K] */, void *stream);
/* Pooled allele counts of R/HoneyBADGER.R:1404-1405: out[k*G+g] = #{c in group k: v[g*ld+c] > 0};
 * v, member [K][N] (0/1) and out are DEVICE pointers. */
int hb_pool_count_pos_dev(const int32_t *v, int64_t ld, int64_t G, int64_t N,
                          const uint8_t *member, int32_t K, int32_t *out, void *stream);

/* ------------------------------------------------------------------ grouping of cells
 * The step before the path: R/HoneyBADGER.R:1122-1136 and :1378-1397 smooth every cell with
 * caTools::runmean, take stats::dist between cells (NA-aware, NA -> 0), build hclust(., "ward.D2")
 * and cut it at k = 1..min.traverse groups.  All pointers are device pointers, gene-major layout.
 * hb_runmean_dev: centred window of k (odd) positions, shrinking at the ends, non-finite values
 *   skipped, NaN for an empty window; hb_runmean_ratio_dev smooths r / n of two count planes.
 * hb_dist_sq_dev: SQUARED Euclidean distance between the N columns of s with R's NA rule (pairs with
 *   a non-finite member are dropped and the sum scaled by G / pairs used; no pair -> 0), D is [N][N].
 * hb_ward_linkage_dev: Ward.D2 by nearest-neighbour chain on D (destroyed).  Merge m joins the
 *   clusters containing leaves merge_a[m] and merge_b[m] at height[m]; N-1 merges in chain order
 *   (children before parents) — sort by height to cut the tree. */
int hb_runmean_dev(const double *x, int64_t ld, int64_t G, int64_t N, int32_t k, double *out, int64_t ldo,
                   void *stream);
int hb_runmean_ratio_dev(const int32_t *r, const int32_t *n, int64_t ld, int64_t G, int64_t N, int32_t k,
                         double *out, int64_t ldo, void *stream);
int hb_dist_sq_dev(const double *s, int64_t ld, int64_t G, int64_t N, double *D, void *stream);
/* A rectangular block of the same distances: rows = the na cells (columns) of sa, columns = the nb cells of sb,
 * D[i * ldd + j], both matrices gene-major with G rows.  Each entry is computed exactly as hb_dist_sq_dev computes
 * it (same operations, same order), so blocks computed on different GPUs assemble into the single-GPU matrix bit
 * for bit — the row-block split of `dist` for cells sharded over ranks (honeybadger_b200/dist.py).  Needs even
 * lda / ldb and 16-byte aligned sa / sb; na, nb >= 2. */
int hb_dist_sq_rect_dev(const double *sa, int64_t lda, int64_t na, const double *sb, int64_t ldb, int64_t nb,
                        int64_t G, double *D, int64_t ldd, void *stream);
int hb_ward_linkage_dev(double *D, int64_t N, int32_t *merge_a, int32_t *merge_b, double *height,
                        void *stream);
/* cutree(hc, k) for each of the nk values in ks, from the merge list hb_ward_linkage_dev wrote (host copies):
 * the N - k lowest merges (stable by height, so children stay before their parents) are applied with
 * union-find and the clusters are numbered by first appearance in cell order, as stats::cutree numbers
 * them (R/HoneyBADGER.R:1136, :1397).  labels is [nk][N], 1-based.  Host only, no device work. */
int hb_cutree_host(const int32_t *merge_a, const int32_t *merge_b, const double *height, int64_t N,
                   const int32_t *ks, int32_t nk, int32_t *labels);

/* Host-buffer forms for the R side (matrices column-major [G x N] as R holds them): the whole
 * `runmean -> dist -> hclust(ward.D2) -> cutree(k)` block of R/HoneyBADGER.R:1122-1136 (window 101) and
 * :1378-1397 (window 31, r.maf / n.sc) in one call.  labels[q*N + c] = cutree(hc, k = ks[q])[c]
 * (1-based, clusters numbered by first appearance as cutree does). */
int hb_group_gexp_host(const double *gexp_norm, int64_t G, int64_t N, int32_t window, const int32_t *ks,
                       int32_t nk, int32_t *labels);
int hb_group_allele_host(const double *r_maf, const double *n_sc, int64_t G, int64_t N, int32_t window,
                         const int32_t *ks, int32_t nk, int32_t *labels);

/* ---- Input construction: the matrices the path consumes, built on the device.
 *
 * setGexpMats (R/HoneyBADGER.R:124-188; function form R/HoneyBADGER_gexp.R:28-90), the arithmetic
 * part: rowMeans filter, scale() of both matrices, gexp.norm = gexp.sc - rowMeans(gexp.ref).  Gene
 * positions (biomaRt) and name intersections stay with the caller; rows of `sc` and `ref` are aligned.
 * R semantics reproduced bit for bit (80-bit LDOUBLE accumulators of rowMeans / colMeans / sum
 * replayed, no FMA contraction); inputs must be finite (no NA).
 *   filter != 0:  keep = (rowMeans(sc) > both & rowMeans(ref) > both) | rowMeans(sc) > test |
 *                 rowMeans(ref) > refthr;  min_mean_test / min_mean_ref NULL = R's defaults
 *                 mean(m[m != 0]) of the respective matrix.
 *   scale != 0:   scale(sc), scale(ref) on the kept rows (centre = colMeans, scale =
 *                 sqrt(sum(v^2) / max(1, n - 1))).
 * Outputs: *Gk kept rows, keep_rows[0..Gk) their 0-based indices (host, capacity G, may be NULL),
 * norm = the Gk x N result.  info (may be NULL): bit 0 / bit 1 set when a row mean fell inside the
 * safety margin of the default test / ref threshold and R's serial mean was replayed exactly. */
int hb_set_gexp_mats_dev(const double *sc, int64_t ld, int64_t G, int64_t N, const double *ref, int64_t ldr,
                         int64_t R, int32_t filter, double min_mean_both, const double *min_mean_test,
                         const double *min_mean_ref, int32_t scale, double *norm /* dev [G][ldo] */,
                         int64_t ldo, int32_t *keep_rows, int64_t *Gk, int32_t *info, void *stream);
/* Host form: sc, ref, norm column-major as R holds them ([G x N], [G x R], [Gk x N]; norm capacity G*N). */
int hb_set_gexp_mats_host(const double *sc, int64_t G, int64_t N, const double *ref, int64_t R, int32_t filter,
                          double min_mean_both, const double *min_mean_test, const double *min_mean_ref,
                          int32_t scale, double *norm, int32_t *keep_rows, int64_t *Gk, int32_t *info);
/* Building blocks (device, gene-major x[g*ld + c]). */
int hb_row_means_dev(const double *x, int64_t ld, int64_t G, int64_t N, double *out /* [G] */, void *stream);
int hb_col_scale_stats_dev(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *keep /* [G] or NULL */,
                           double *center /* [N] */, double *scl /* [N] or NULL */, void *stream);
int hb_gexp_apply_dev(const double *x, int64_t ld, int64_t N, const int32_t *rows /* [Gk] or NULL */, int64_t Gk,
                      const double *center, const double *scl, const double *refmean /* [Gk] */, double *out,
                      int64_t ldo, void *stream);
/* R's mean(x[x != 0]) into *mean (host).  exact = 0: parallel double-double sum, *margin (host, may be
 * NULL) bounds the distance to R's serially rounded value; exact = 1: R's two-pass loop replayed. */
int hb_mean_nonzero_dev(const double *x, int64_t ld, int64_t G, int64_t N, int32_t exact, double *mean,
                        double *margin, void *stream);

/* setAlleleMats (R/HoneyBADGER.R:486-614; function form R/HoneyBADGER_allele.R:23-120): in-silico bulk
 * l = rowSums(r > 0), n.bulk = rowSums(n.sc > 0) unless both are given (host [G]); filter: E = l / n.bulk
 * in (thr, 1 - thr) and coverage in >= min_cell cells; lesser-allele flip r.maf = E <= 0.5 ? r : n.sc - r
 * (0 where E is NA), l.maf likewise.  Integer, exact.  Outputs: r_maf, n_sc ([Gk] rows; n_sc may be
 * NULL), l, n_bulk, l_maf (host [Gk], capacity G, may be NULL), keep_rows, *Gk, and *n_het (may be
 * NULL) = SNPs that passed the heterozygosity filter, the number the reference prints (:527). */
int hb_set_allele_mats_dev(const int32_t *r, const int32_t *n_sc_init, int64_t ld, int64_t G, int64_t N,
                           const double *l_init, const double *n_bulk_init, int32_t filter,
                           double het_deviance_threshold, int32_t min_cell, int32_t *r_maf /* dev [G][ldo] */,
                           int32_t *n_sc /* dev [G][ldo] or NULL */, int64_t ldo, double *l, double *n_bulk,
                           double *l_maf, int32_t *keep_rows, int64_t *Gk, int64_t *n_het, void *stream);
/* Host form: r, n_sc_init, r_maf, n_sc column-major int32 as R holds integer matrices. */
int hb_set_allele_mats_host(const int32_t *r, const int32_t *n_sc_init, int64_t G, int64_t N,
                            const double *l_init, const double *n_bulk_init, int32_t filter,
                            double het_deviance_threshold, int32_t min_cell, int32_t *r_maf, int32_t *n_sc,
                            double *l, double *n_bulk, double *l_maf, int32_t *keep_rows, int64_t *Gk,
                            int64_t *n_het);

/* Deterministic retest of one candidate region for the expression model: the exact marginal posterior
 * of inst/bug/expressionModel.bug that HoneyBADGER$calcGexpCnvProb
 * (R/HoneyBADGER.R:374-470, called from :1220) estimates by JAGS sampling — statistical stand-in for
 * the MCMC, not bit parity.  x is the resident gene-major matrix, rows[nrows] (device) the region's
 * gene rows, sigma0[N] (device or NULL = 1) the per-cell dnorm PRECISION exactly as the reference
 * passes it (expressionModel.bug:8), mag0 the deviation.  Outputs per cell (device): p_amp =
 * mean(S dd), p_del = mean(S (1-dd)), mu0 = column mean of the region (or NULL).
 * phase 0: everything on one GPU.  Cells sharded over ranks: phase 1 computes the per-cell weights
 * and this rank's evidence D_local (device scalar); all-reduce D over the ranks; phase 2 applies
 * the shared direction posterior from D_total (device scalar). */
int hb_retest_gexp_dev(const double *x, int64_t ld, int64_t G, int64_t N, const int64_t *rows,
                       int64_t nrows, double mag0, const double *sigma0, double *p_amp, double *p_del,
                       double *mu0, double *D_local, const double *D_total, int32_t phase, void *stream);

/* Host-buffer form: gexp_rows column-major [n x N] = gexp.norm[bound.genes.new, ] (R/HoneyBADGER.R:1220),
 * outputs of length N; mu0 may be NULL. */
int hb_retest_gexp_host(const double *gexp_rows, int64_t n, int64_t N, double mag0, const double *sigma0,
                        double *p_amp, double *p_del, double *mu0);

/* Allele counterpart (stand-in for the JAGS snpModel of calcAlleleCnvProb, R/HoneyBADGER.R:907-1042):
 * P(deletion) per cell from the binomial likelihood ratio pd vs pn summed over the region's SNPs. */
int hb_retest_allele_lr_dev(const int32_t *r_maf, const int32_t *n_sc, int64_t ld, int64_t G, int64_t N,
                            const int64_t *rows, int64_t nrows, double pd, double pn, double *p_del,
                            void *stream);

/* The allele model itself: exact marginal posterior P(S_k = 1 | data) of inst/bug/snpModel.bug, which
 * HoneyBADGER$calcAlleleCnvProb (R/HoneyBADGER.R:907-1042, caller :1486 with pe = pd, mono = 0.7) samples
 * with JAGS.  Per cell every latent variable (S, alpha, per-gene b and d, per-SNP h) is integrated out —
 * h by a 256-node composite Gauss-Legendre rule over the truncated normal — and the per-SNP allele
 * indicator ma is replaced by its bulk posterior P(ma = 1 | l.maf, n.bulk) (the single approximation).
 * Statistical stand-in for the MCMC, not bit parity.  rows[nrows] (device) are the region's SNP rows
 * grouped by gene, gene_end[nrows] (host, NULL = one gene per SNP) marks each gene's last SNP, l_maf /
 * n_bulk (host [nrows]) the bulk counts in the same order.  p_del[N] on the device; logodds != 0 stores
 * log P(data | S = 1) - log P(data | S = 0) instead of the probability. */
int hb_retest_allele_model_dev(const int32_t *r_maf, const int32_t *n_sc, int64_t ld, int64_t G, int64_t N,
                               const int64_t *rows, int64_t nrows, const uint8_t *gene_end,
                               const double *l_maf, const double *n_bulk, double pseudo, double mono,
                               int32_t logodds, double *p_del, void *stream);
/* The combined model (inst/bug/combinedModel.bug; HoneyBADGER$calcCombCnvProb, R/HoneyBADGER.R:1620-1782,
 * caller :1866): expression and allele evidence share S_k and the direction dd.  Its exact marginal is the
 * expression retest with the deletion hypothesis of every cell shifted by that cell's allele log-odds
 * log P(alleles | deleted) - log P(alleles | neutral) = hb_retest_allele_model_dev(..., logodds = 1, ...):
 * pass that vector as del_logodds (device [N]; NULL = hb_retest_gexp_dev).  Phases as hb_retest_gexp_dev. */
int hb_retest_comb_dev(const double *x, int64_t ld, int64_t G, int64_t N, const int64_t *rows,
                       int64_t nrows, double mag0, const double *sigma0, const double *del_logodds,
                       double *p_amp, double *p_del, double *mu0, double *D_local, const double *D_total,
                       int32_t phase, void *stream);
/* Host form: the region's [n x N] column-major integer sub-matrices, gene[n] factor codes (NULL = one gene
 * per SNP; grouped by first appearance like unique(geneFactor), R/HoneyBADGER.R:963). */
int hb_retest_allele_model_host(const int32_t *r_rows, const int32_t *n_rows, int64_t n, int64_t N,
                                const double *l_maf, const double *n_bulk, const int32_t *gene, double pseudo,
                                double mono, double *p_del);

/* ------------------------------------------------------------------ a few LONG chains
 * HiddenMarkov::Viterbi on precomputed log-emissions, parallel ALONG the chain (the pooled rounds of
 * R/HoneyBADGER.R:1150-1151 and :1409-1410 run K <= a few dozen chains of G positions; one thread per
 * chain would leave the GPU idle).  Blocks of positions get exact integer max-plus transfer matrices
 * (inside one binade the rounded recursion is an integer recursion), block entries are predicted from
 * them, and every block is then run with the literal fp64 recursion from its entry; a chain's result is
 * accepted only when every block's literal exit is bit-identical to the next block's entry, otherwise
 * the rest of the chain is finished sequentially — the states are the sequential recursion's in all cases.
 * logb: device [K][T][S] (chain-major; the reference's own dnorm / dbinom values), S = 2 or 3,
 * Pi [S*S] row-major, delta [S], block_len 0 = chosen from T, y: device [K][T] 1-based.
 * No synchronisation; hb_status_dev reports HiddenMarkov's underflow error. */
int hb_viterbi_long_logb_dev(const double *logb, int32_t S, int64_t T, int32_t K, const double *Pi,
                             const double *delta, int32_t block_len, int32_t *y, void *stream);
/* State posteriors of the same chains, parallel along the chain: gamma_t(k) = P(state_t = k | all x), the
 * semantics of HiddenMarkov::forwardback followed by exp(logalpha + logbeta - LL) (the reference
 * never runs it on the pooled chains — north_star capability).  Linear domain with per-position
 * normalisation on emissions exp(logb_t - max_k logb_t), one S x S product per block of positions serving
 * both sweeps; agreement with the oracle 1e-9 (the oracle's own log-space rounding grows with T).
 * gamma: device [K][T][S]; ll: device [K] log-likelihoods or NULL.  No synchronisation. */
int hb_forwardback_long_logb_dev(const double *logb, int32_t S, int64_t T, int32_t K, const double *Pi,
                                 const double *delta, int32_t block_len, double *gamma, double *ll,
                                 void *stream);
/* Counters of the last long-chain launch (waits for `stream`): out[0] chains, out[1] blocks per chain,
 * out[2] chains finished on the sequential path, out[3] repair rounds used (max over chains), out[4]
 * blocks the sequential walkers ran literally (entry outside the predicted binade, first blocks). */
int hb_longchain_stats_dev(void *stream, int64_t *out /* [5] */);
/* Where the long-chain path is used.  0 (default): the pooled allele round, and hb_hmm_norm_host /
 * hb_hmm_binom_host calls of the reference's own shape (<= 32 chains, >= 1024 positions, one segment —
 * Viterbi(dthmm(x, ...)) on a pooled vector; states, posteriors and LL); 1: nowhere (one thread per chain);
 * 2: as 0 plus the pooled expression round.  Same states in every mode, posteriors within 1e-9. */
int hb_set_longchain_mode(int32_t mode);
/* Which kernel builds the Ward linkage (hb_ward_linkage_dev and the hb_group_* entry points).  0 (default):
 * from 256 cells up a thread-block cluster of up to 16 SMs (column slices per CTA, winners exchanged
 * through distributed shared memory), one block below that; 1: always one block; 2: always the cluster.
 * Merges, their order and heights are bit-identical in every mode.  Per calling thread. */
int hb_set_ward_mode(int32_t mode);
/* Which kernel computes the distances.  0 (default): slabs of positions staged with cp.async three deep and a
 * pre-pass that marks the slabs holding non-finite values; 1: the plain staged kernel.  Bit-identical
 * results.  Per calling thread. */
int hb_set_dist_mode(int32_t mode);

/* One recursion round of the split-and-retest driver on DEVICE-RESIDENT matrices (gene-major,
 * uploaded once): `member` [K][N] (0/1, host) names the K candidate cell sets of the round, the
 * small results come back to host buffers: y [K][G] (1-based states), pooled vectors and their sd.
 * The [G x N] matrices never leave the GPU between rounds.  Replaces the `lapply(heights, ...)`
 * bodies at R/HoneyBADGER.R:1135-1161 and :1395-1417. */
int hb_gexp_pooled_viterbi_dev(const double *x /* device x[g*ld+c] */, int64_t ld, int64_t G,
                               int64_t N, const uint8_t *member /* host */, int32_t K,
                               const double *mean, const double *Pi, const double *delta,
                               int32_t *y /* host [K][G] */, double *pooled_x /* host or NULL */,
                               double *pooled_sd /* host [K] or NULL */, void *stream);
int hb_allele_pooled_viterbi_dev(const int32_t *r_maf /* device */, const int32_t *n_sc, int64_t ld,
                                 int64_t G, int64_t N, const uint8_t *member /* host */, int32_t K,
                                 const double *prob, const double *Pi, const double *delta,
                                 int32_t *y /* host [K][G] */, int32_t *mafl /* host or NULL */,
                                 int32_t *sizel /* host or NULL */, void *stream);

/* Host-buffer convenience for the four reference call sites (K pooled chains of one round):
 * gexp_norm column-major [G x N]; member[K][N] (0/1).  Outputs y[K][G] (1-based), pooled x and sd. */
int hb_gexp_pooled_viterbi_host(const double *gexp_norm, int64_t G, int64_t N,
                                const uint8_t *member, int32_t K, const double *mean,
                                const double *Pi, const double *delta, int32_t *y /* [K][G] */,
                                double *pooled_x /* [K][G] or NULL */, double *pooled_sd /* [K] */);
int hb_allele_pooled_viterbi_host(const double *r_maf, const double *n_sc, int64_t G, int64_t N,
                                  const uint8_t *member, int32_t K, const double *prob,
                                  const double *Pi, const double *delta, int32_t *y /* [K][G] */,
                                  int32_t *mafl /* [K][G] or NULL */,
                                  int32_t *sizel /* [K][G] or NULL */);

/* ------------------------------------------------------------------ device-resident handle
 * The state the reference keeps across its recursive calls is the matrices themselves (RefClass fields,
 * R/HoneyBADGER.R:1098-1102), and every recursion level copies sub-matrices `gexp.norm[, g1]`
 * (:1304-1309, allele :1559-1575).  With a handle the matrices cross PCIe ONCE: hb_upload_* turns R's
 * column-major doubles into the resident gene-major layout and returns an opaque pointer (the .Call shim wraps
 * it in an EXTPTRSXP whose finalizer is hb_free_handle and keeps it in a RefClass field, r/hb_shim.c,
 * r/hb_b200.R); each later call names the candidate genes / cells of the level by 0-based index sets (host
 * int32 arrays, NULL = all) and only those indices, the [K x ncols] membership and the small results move.
 * Row / cell order of the results is the order of the index sets.  All calls run on the legacy stream and
 * return finished host results. */
typedef struct hb_handle hb_handle;
int hb_upload_gexp_host(const double *gexp_norm /* [G x N] column-major */, int64_t G, int64_t N, hb_handle **out);
int hb_upload_counts_host(const double *r_maf, const double *n_sc /* [G x N] column-major, R doubles */,
                          int64_t G, int64_t N, hb_handle **out);
int hb_free_handle(hb_handle *h);
/* kind 1 = expression, 2 = allele counts; h2d_bytes = bytes the uploads moved host -> device (the later calls
 * add only index sets and memberships).  Any output pointer may be NULL. */
int hb_handle_info(const hb_handle *h, int32_t *kind, int64_t *G, int64_t *N, int64_t *h2d_bytes);
/* One `lapply(heights, ...)` round (R/HoneyBADGER.R:1135-1161 / :1395-1417) on the sub-matrix
 * [rows x cols] of the resident data: member [K][ncols] (0/1) over the SELECTED cells, outputs as
 * hb_gexp_pooled_viterbi_host / hb_allele_pooled_viterbi_host with G = nrows. */
int hb_gexp_round_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                         const uint8_t *member, int32_t K, const double *mean, const double *Pi,
                         const double *delta, int32_t *y /* [K][nrows] */, double *pooled_x, double *pooled_sd);
int hb_allele_round_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                           const uint8_t *member, int32_t K, const double *prob, const double *Pi,
                           const double *delta, int32_t *y /* [K][nrows] */, int32_t *mafl, int32_t *sizel);
/* runmean -> dist -> hclust(ward.D2) -> cutree(k = ks[q]) on the same sub-matrix (R/HoneyBADGER.R:1122-1136 with
 * window 101; :1378-1397 with window 31 on r.maf / n.sc): labels[q * ncols + c], 1-based. */
int hb_group_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                    int32_t window, const int32_t *ks, int32_t nk, int32_t *labels);
/* The retests of one candidate region (rows) over the level's cells (cols): hb_retest_gexp_host /
 * hb_retest_allele_model_host on the resident data.  sigma0 (host [ncols] or NULL), l_maf / n_bulk (host
 * [nrows]), gene_end (host [nrows] or NULL = one gene per SNP; rows must already be grouped by gene). */
int hb_retest_gexp_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                          double mag0, const double *sigma0, double *p_amp, double *p_del, double *mu0);
int hb_retest_allele_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                            const double *l_maf, const double *n_bulk, const uint8_t *gene_end, double pseudo,
                            double mono, double *p_del);

#ifdef __cplusplus
}
#endif
#endif /* HB_B200_H */
