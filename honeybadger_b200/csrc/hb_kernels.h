// hb_kernels.h — launcher declarations shared by hb_kernels.cu and hb_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HB_BLOCK 128

namespace hb {

struct Norm3Args;
struct Norm3Const;
struct Binom2Args;
struct LcArgs;
struct FbrArgs;

cudaError_t launch_longchain(const LcArgs &a, cudaStream_t st);
cudaError_t launch_longchain_fb(const LcArgs &a, cudaStream_t st);
cudaError_t launch_binom2_emit_cm(const Binom2Args &a, const int32_t *x, const int32_t *n, int64_t count,
                                  double2 *out, cudaStream_t st);
cudaError_t launch_norm3_emit_cm(const Norm3Const &c, const double *x, int64_t G, int32_t K, const double *sd,
                                 const double *inv_sd, const double *log_sd, double *out, cudaStream_t st);

cudaError_t launch_norm3(const Norm3Args &a, bool sym, bool vit, bool fb, cudaStream_t st);
cudaError_t launch_gather_rc(const void *x, int32_t elem, int64_t ld, const int32_t *rows, int64_t nrows,
                             const int32_t *cols, int64_t ncols, void *out, int64_t ldo, cudaStream_t st);
cudaError_t launch_norm3_resident(const FbrArgs &a, int nw, size_t smem, cudaStream_t st);
cudaError_t launch_norm3_fast(const Norm3Args &a, bool vit, bool fb, cudaStream_t st);
cudaError_t launch_pack_counts(const int32_t *x, const int32_t *n, int64_t count, uint32_t *out,
                               int32_t *overflow, cudaStream_t st);
cudaError_t launch_binom2_emit(const Binom2Args &a, double2 *pre_lb, cudaStream_t st);
cudaError_t launch_binom2_fast(const Binom2Args &a, bool vit, bool fb, cudaStream_t st);
cudaError_t launch_binom2(const Binom2Args &a, bool vit, bool fb, cudaStream_t st);
cudaError_t launch_sd_prep(const double *sd, int64_t n, double *inv_sd, double *log_sd,
                           cudaStream_t st);
cudaError_t launch_chain_sd(const double *x, int64_t ld, int64_t B, const int64_t *seg_off_dev,
                            int32_t nseg, double *sd, cudaStream_t st);
cudaError_t launch_transpose_f64(const double *src, int64_t G, int64_t N, double *dst, int64_t ld,
                                 cudaStream_t st);
cudaError_t launch_transpose_f64_i32(const double *src, int64_t G, int64_t N, int32_t *dst,
                                     int64_t ld, cudaStream_t st);
cudaError_t launch_transpose_i32(const int32_t *src, int64_t G, int64_t N, int32_t *dst,
                                 int64_t ld, cudaStream_t st);
cudaError_t launch_transpose_u8_back(const uint8_t *src, int64_t ld, int64_t G, int64_t N,
                                     int32_t *dst, cudaStream_t st);
cudaError_t launch_transpose_u8_back_u8(const uint8_t *src, int64_t ld, int64_t G, int64_t N,
                                        uint8_t *dst, cudaStream_t st);
cudaError_t launch_transpose_f64_back(const double *src, int64_t ld, int64_t G, int64_t N,
                                      double *dst, cudaStream_t st);
cudaError_t launch_transpose_i32_back(const int32_t *src, int64_t ld, int64_t G, int64_t N,
                                      int32_t *dst, cudaStream_t st);
cudaError_t launch_tile(const void *src, int64_t ld, int64_t T, int64_t B, int elem, void *dst,
                        bool to_tiled, cudaStream_t st);
cudaError_t launch_retest_allele(const int32_t *r, const int32_t *n, int64_t ld, int64_t N,
                                 const int64_t *rows, int64_t nrows, double pd, double pn, double *p_del,
                                 cudaStream_t st);
cudaError_t launch_retest_allele_model(const int32_t *r, const int32_t *n, int64_t ld, int64_t N,
                                       const int64_t *rows, int64_t nrows, const uint8_t *gene_end,
                                       const double *lq1, const double *lq0, const double *tab, int32_t nq,
                                       double pseudo, double mono, int32_t logodds, double *p_del,
                                       cudaStream_t st);
cudaError_t launch_retest_gexp_a(const double *x, int64_t ld, int64_t N, const int64_t *rows,
                                 int64_t nrows, double mag0, const double *sigma0, const double *del_off,
                                 double *w_amp, double *w_del, double *mu0, double *partial, double *D_local,
                                 cudaStream_t st);
cudaError_t launch_retest_gexp_b(int64_t N, const double *partial, const double *D_total, double *p_amp,
                                 double *p_del, cudaStream_t st);
cudaError_t launch_runmean(const double *x, const int32_t *r, const int32_t *n, int64_t ld, int64_t G,
                           int64_t N, int32_t k, double *out, int64_t ldo, double *tmp, cudaStream_t st);
size_t dist_scratch_bytes(int64_t G, int64_t N);
cudaError_t launch_dist_sq(const double *S, int64_t ld, int64_t G, int64_t N, double *D, void *scratch, cudaStream_t st);
cudaError_t launch_dist_sq_rect(const double *Sa, int64_t lda, int64_t Na, const double *Sb, int64_t ldb, int64_t Nb,
                                int64_t G, double *D, int64_t ldd, void *scratch, cudaStream_t st);
size_t ward_scratch_bytes(int64_t N);
cudaError_t launch_ward(double *D, int64_t N, void *work, int32_t *ma, int32_t *mb, double *height, int mode,
                        cudaStream_t st);
cudaError_t launch_pool_mean(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                             const int32_t *gsize, int32_t K, double *out, void *scratch, void *fast,
                             const int32_t *order, cudaStream_t st);
size_t pool_accum_scratch_bytes(int64_t G, int64_t N, int32_t K);
cudaError_t launch_pool_accum_prepare(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                                      int32_t K, void *scratch, cudaStream_t st);
cudaError_t launch_pool_accum_fast(int64_t G, int64_t N, int32_t K, const int32_t *order, const void *mean80,
                                   void *state, const void *scratch, cudaStream_t st);
cudaError_t launch_pool_accum(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                              int32_t K, const void *mean80, void *state, cudaStream_t st);
cudaError_t launch_pool_div(const void *state, const int32_t *n, int64_t G, int32_t K, void *mean80,
                            cudaStream_t st);
cudaError_t launch_pool_finish(const void *mean80, const void *state, const int32_t *n, int64_t G,
                               int32_t K, double *out, cudaStream_t st);
cudaError_t launch_pooled_sd(const double *pooled, int64_t G, int32_t K, double *sd,
                             cudaStream_t st);
cudaError_t launch_pool_count(const int32_t *v, int64_t ld, int64_t G, int64_t N,
                              const uint32_t *cellmask, int32_t K, int32_t *out, cudaStream_t st);
cudaError_t launch_build_cellmask(const uint8_t *member, int64_t N, int32_t K, uint32_t *mask,
                                  cudaStream_t st);

cudaError_t launch_row_means(const double *x, int64_t ld, int64_t G, int64_t N, double *out, cudaStream_t st);
cudaError_t launch_col_stats(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *keep,
                             double *center, double *scl, cudaStream_t st);
cudaError_t launch_gexp_apply(const double *x, int64_t ld, int64_t N, const int32_t *rows, int64_t Gk,
                              const double *center, const double *scl, const double *refmean, double *out,
                              int64_t ldo, cudaStream_t st);
cudaError_t launch_nz_sum(const double *x, int64_t ld, int64_t G, int64_t N, int nblocks, double *part_hi,
                          double *part_lo, unsigned long long *part_n, double *part_max, cudaStream_t st);
cudaError_t launch_nz_mean_seq(const double *x, int64_t ld, int64_t G, int64_t N, unsigned long long n,
                               double *out, cudaStream_t st);
cudaError_t launch_allele_flip(const int32_t *r, const int32_t *n, int64_t ld, int64_t N, const int32_t *rows,
                               int64_t Gk, const uint8_t *code, int32_t *out_r, int32_t *out_n, int64_t ldo,
                               cudaStream_t st);

} // namespace hb
