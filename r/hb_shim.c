/*
 * hb_shim.c — `.Call` shim between R and libhb_b200.so (include/hb_b200.h).
 *
 * NOT built in this repository: the build container and the GPU boxes have no R (no R.h /
 * Rinternals.h), so this file is the integration a HoneyBADGER maintainer adds to the package's
 * src/ (see INTEGRATION.md).  It is deliberately thin: coerce, allocate, call, convert the error.
 *
 *   R CMD SHLIB hb_shim.c -I<repo>/include -L<repo>/honeybadger_b200 -lhb_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "hb_b200.h"

static void hb_raise(int rc)
{
    if (rc == HB_OK) return;
    if (rc == HB_ERR_UNDERFLOW) Rf_error("Problems With Underflow"); /* HiddenMarkov::Viterbi's text */
    Rf_error("libhb_b200: %s (code %d)", hb_last_error(), rc);
}

/* y <- .Call(C_hb_viterbi_norm, x, mean, sd, Pi, delta)
 * replaces  HiddenMarkov::Viterbi(HiddenMarkov::dthmm(x, Pi, delta, "norm", list(mean=, sd=)))
 * x: numeric vector or [T x B] matrix (one chain per column); Pi as the R matrix (column-major). */
SEXP C_hb_viterbi_norm(SEXP x, SEXP mean, SEXP sd, SEXP Pi, SEXP delta)
{
    x = PROTECT(Rf_coerceVector(x, REALSXP));
    const int is_mat = Rf_isMatrix(x);
    const int64_t T = is_mat ? Rf_nrows(x) : XLENGTH(x), B = is_mat ? Rf_ncols(x) : 1;
    double pi_rm[9];
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) pi_rm[j * 3 + k] = REAL(Pi)[k * 3 + j]; /* R is column-major */
    SEXP sdv = PROTECT(Rf_allocVector(REALSXP, B));
    for (int64_t b = 0; b < B; b++) REAL(sdv)[b] = REAL(sd)[XLENGTH(sd) == B ? b : 0];
    SEXP y = PROTECT(is_mat ? Rf_allocMatrix(INTSXP, (int)T, (int)B) : Rf_allocVector(INTSXP, T));
    int rc = hb_hmm_norm_host(REAL(x), T, B, NULL, 1, REAL(mean), REAL(sdv), 0, pi_rm, REAL(delta),
                              HB_WANT_VITERBI, INTEGER(y), NULL, NULL);
    UNPROTECT(3);
    hb_raise(rc);
    return y;
}

/* y <- .Call(C_hb_viterbi_binom, x, size, prob, Pi, delta)
 * replaces  Viterbi(dthmm(x, Pi, delta, "binom", list(prob=), list(size=), discrete=TRUE)) */
SEXP C_hb_viterbi_binom(SEXP x, SEXP size, SEXP prob, SEXP Pi, SEXP delta)
{
    x = PROTECT(Rf_coerceVector(x, INTSXP));
    size = PROTECT(Rf_coerceVector(size, INTSXP));
    const int is_mat = Rf_isMatrix(x);
    const int64_t T = is_mat ? Rf_nrows(x) : XLENGTH(x), B = is_mat ? Rf_ncols(x) : 1;
    double pi_rm[4] = {REAL(Pi)[0], REAL(Pi)[2], REAL(Pi)[1], REAL(Pi)[3]};
    SEXP y = PROTECT(is_mat ? Rf_allocMatrix(INTSXP, (int)T, (int)B) : Rf_allocVector(INTSXP, T));
    int rc = hb_hmm_binom_host(INTEGER(x), INTEGER(size), T, B, NULL, 1, REAL(prob), 0.0, pi_rm,
                               REAL(delta), HB_WANT_VITERBI, INTEGER(y), NULL, NULL);
    UNPROTECT(3);
    hb_raise(rc);
    return y;
}

/* One whole `lapply(heights, ...)` round: pooled mean + sd + Viterbi for K cell sets at once.
 * member: logical/integer [N x K] matrix (ct == group columns).  Returns list(y = [G x K] integer). */
SEXP C_hb_gexp_pooled_viterbi(SEXP gexp_norm, SEXP member, SEXP mean, SEXP Pi, SEXP delta)
{
    gexp_norm = PROTECT(Rf_coerceVector(gexp_norm, REALSXP));
    member = PROTECT(Rf_coerceVector(member, INTSXP));
    const int64_t G = Rf_nrows(gexp_norm), N = Rf_ncols(gexp_norm);
    const int K = Rf_ncols(member);
    unsigned char *mem = (unsigned char *)R_alloc((size_t)K * N, 1);
    for (int k = 0; k < K; k++)
        for (int64_t c = 0; c < N; c++) mem[(size_t)k * N + c] = INTEGER(member)[(size_t)k * N + c] != 0;
    double pi_rm[9];
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) pi_rm[j * 3 + k] = REAL(Pi)[k * 3 + j];
    SEXP y = PROTECT(Rf_allocMatrix(INTSXP, (int)G, K)); /* [K][G] row-major == [G x K] column-major */
    int rc = hb_gexp_pooled_viterbi_host(REAL(gexp_norm), G, N, mem, K, REAL(mean), pi_rm, REAL(delta),
                                         INTEGER(y), NULL, NULL);
    UNPROTECT(3);
    hb_raise(rc);
    return y;
}

SEXP C_hb_allele_pooled_viterbi(SEXP r_maf, SEXP n_sc, SEXP member, SEXP prob, SEXP Pi, SEXP delta)
{
    r_maf = PROTECT(Rf_coerceVector(r_maf, REALSXP));
    n_sc = PROTECT(Rf_coerceVector(n_sc, REALSXP));
    member = PROTECT(Rf_coerceVector(member, INTSXP));
    const int64_t G = Rf_nrows(r_maf), N = Rf_ncols(r_maf);
    const int K = Rf_ncols(member);
    unsigned char *mem = (unsigned char *)R_alloc((size_t)K * N, 1);
    for (size_t i = 0; i < (size_t)K * N; i++) mem[i] = INTEGER(member)[i] != 0;
    double pi_rm[4] = {REAL(Pi)[0], REAL(Pi)[2], REAL(Pi)[1], REAL(Pi)[3]};
    SEXP y = PROTECT(Rf_allocMatrix(INTSXP, (int)G, K));
    int rc = hb_allele_pooled_viterbi_host(REAL(r_maf), REAL(n_sc), G, N, mem, K, REAL(prob), pi_rm,
                                           REAL(delta), INTEGER(y), NULL, NULL);
    UNPROTECT(4);
    hb_raise(rc);
    return y;
}

/* ct <- .Call(C_hb_group_gexp, gexp.norm, 101L, heights): replaces, at R/HoneyBADGER.R:1122-1136,
 *   mat.smooth <- apply(gexp.norm, 2, caTools::runmean, k); d <- dist(t(mat.smooth)); d[is.na(d)] <- 0
 *   hc <- hclust(d, method = "ward.D2");  lapply(heights, function(h) cutree(hc, k = h))
 * Returns an integer [N x length(heights)] matrix: column q is cutree(hc, k = heights[q]). */
SEXP C_hb_group_gexp(SEXP gexp_norm, SEXP window, SEXP heights)
{
    gexp_norm = PROTECT(Rf_coerceVector(gexp_norm, REALSXP));
    heights = PROTECT(Rf_coerceVector(heights, INTSXP));
    const int64_t G = Rf_nrows(gexp_norm), N = Rf_ncols(gexp_norm);
    const int nk = (int)XLENGTH(heights);
    SEXP lab = PROTECT(Rf_allocMatrix(INTSXP, (int)N, nk)); /* [nk][N] row-major == [N x nk] column-major */
    int rc = hb_group_gexp_host(REAL(gexp_norm), G, N, Rf_asInteger(window), INTEGER(heights), nk, INTEGER(lab));
    UNPROTECT(3);
    hb_raise(rc);
    return lab;
}

/* The allele form (R/HoneyBADGER.R:1378-1397): r.maf / n.sc smoothed with k = 31. */
SEXP C_hb_group_allele(SEXP r_maf, SEXP n_sc, SEXP window, SEXP heights)
{
    r_maf = PROTECT(Rf_coerceVector(r_maf, REALSXP));
    n_sc = PROTECT(Rf_coerceVector(n_sc, REALSXP));
    heights = PROTECT(Rf_coerceVector(heights, INTSXP));
    const int64_t G = Rf_nrows(r_maf), N = Rf_ncols(r_maf);
    const int nk = (int)XLENGTH(heights);
    SEXP lab = PROTECT(Rf_allocMatrix(INTSXP, (int)N, nk));
    int rc = hb_group_allele_host(REAL(r_maf), REAL(n_sc), G, N, Rf_asInteger(window), INTEGER(heights), nk,
                                  INTEGER(lab));
    UNPROTECT(4);
    hb_raise(rc);
    return lab;
}

/* prob <- .Call(C_hb_retest_gexp, gexp.norm[bound.genes.new, ], m, sigma0): the exact marginal of
 * inst/bug/expressionModel.bug instead of the JAGS run of calcGexpCnvProb (R/HoneyBADGER.R:374-470).
 * Returns list(amp, del, mu0), the three vectors calcGexpCnvProb returns. */
SEXP C_hb_retest_gexp(SEXP gexp_rows, SEXP mag0, SEXP sigma0)
{
    gexp_rows = PROTECT(Rf_coerceVector(gexp_rows, REALSXP));
    const int64_t n = Rf_nrows(gexp_rows), N = Rf_ncols(gexp_rows);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    for (int i = 0; i < 3; i++) SET_VECTOR_ELT(out, i, Rf_allocVector(REALSXP, N));
    int rc = hb_retest_gexp_host(REAL(gexp_rows), n, N, Rf_asReal(mag0),
                                 Rf_isNull(sigma0) ? NULL : REAL(sigma0), REAL(VECTOR_ELT(out, 0)),
                                 REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)));
    UNPROTECT(2);
    hb_raise(rc);
    return out;
}

/* res <- .Call(C_hb_set_gexp_mats, gexp.sc, gexp.ref, filter, minMeanBoth, minMeanTest, minMeanRef, scale)
 * replaces the arithmetic of setGexpMats, R/HoneyBADGER.R:139-161 (rowMeans filter, scale(), gexp.sc -
 * rowMeans(gexp.ref)); rows of the two matrices already intersected (:135-137).  minMeanTest / minMeanRef
 * NULL = the defaults mean(m[m != 0]).  Returns list(gexp.norm [Gk x N], keep = 1-based kept rows). */
SEXP C_hb_set_gexp_mats(SEXP sc, SEXP ref, SEXP filter, SEXP min_both, SEXP min_test, SEXP min_ref, SEXP scale)
{
    sc = PROTECT(Rf_coerceVector(sc, REALSXP));
    ref = PROTECT(Rf_coerceVector(ref, REALSXP));
    const int64_t G = Rf_nrows(sc), N = Rf_ncols(sc), R = Rf_isMatrix(ref) ? Rf_ncols(ref) : 1;
    double mt = Rf_isNull(min_test) ? 0 : Rf_asReal(min_test), mr = Rf_isNull(min_ref) ? 0 : Rf_asReal(min_ref);
    SEXP full = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)G * N));
    SEXP keep = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)G));
    int64_t Gk = 0;
    int rc = hb_set_gexp_mats_host(REAL(sc), G, N, REAL(ref), R, Rf_asLogical(filter), Rf_asReal(min_both),
                                   Rf_isNull(min_test) ? NULL : &mt, Rf_isNull(min_ref) ? NULL : &mr,
                                   Rf_asLogical(scale), REAL(full), INTEGER(keep), &Gk, NULL);
    if (rc != 0) {
        UNPROTECT(4);
        hb_raise(rc);
    }
    SEXP norm = PROTECT(Rf_allocMatrix(REALSXP, (int)Gk, (int)N)); /* the library wrote Gk x N contiguously */
    memcpy(REAL(norm), REAL(full), (size_t)Gk * N * sizeof(double));
    SEXP keep1 = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)Gk));
    for (int64_t k = 0; k < Gk; k++) INTEGER(keep1)[k] = INTEGER(keep)[k] + 1;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, norm);
    SET_VECTOR_ELT(out, 1, keep1);
    UNPROTECT(7);
    return out;
}

/* res <- .Call(C_hb_set_allele_mats, r.init, n.sc.init, l.init, n.bulk.init, filter, het.deviance.threshold,
 * min.cell) replaces R/HoneyBADGER.R:493-595 (in-silico bulk, both filters, the nested lapply of the
 * lesser-allele flip).  Returns list(r.maf, n.sc, l, n.bulk, l.maf, keep (1-based), n.het). */
SEXP C_hb_set_allele_mats(SEXP r, SEXP n_sc, SEXP l_init, SEXP n_bulk_init, SEXP filter, SEXP thr, SEXP min_cell)
{
    r = PROTECT(Rf_coerceVector(r, INTSXP));
    n_sc = PROTECT(Rf_coerceVector(n_sc, INTSXP));
    const int64_t G = Rf_nrows(r), N = Rf_ncols(r);
    const int given = !Rf_isNull(l_init) && !Rf_isNull(n_bulk_init);
    SEXP li = PROTECT(given ? Rf_coerceVector(l_init, REALSXP) : R_NilValue);
    SEXP ni = PROTECT(given ? Rf_coerceVector(n_bulk_init, REALSXP) : R_NilValue);
    int32_t *o_r = (int32_t *)R_alloc((size_t)G * N, sizeof(int32_t));
    int32_t *o_n = (int32_t *)R_alloc((size_t)G * N, sizeof(int32_t));
    double *l = (double *)R_alloc((size_t)G * 3, sizeof(double)), *nb = l + G, *lm = nb + G;
    int32_t *keep = (int32_t *)R_alloc((size_t)G, sizeof(int32_t));
    int64_t Gk = 0, nhet = 0;
    int rc = hb_set_allele_mats_host(INTEGER(r), INTEGER(n_sc), G, N, given ? REAL(li) : NULL,
                                     given ? REAL(ni) : NULL, Rf_asLogical(filter), Rf_asReal(thr),
                                     Rf_asInteger(min_cell), o_r, o_n, l, nb, lm, keep, &Gk, &nhet);
    if (rc != 0) {
        UNPROTECT(4);
        hb_raise(rc);
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
    SET_VECTOR_ELT(out, 0, Rf_allocMatrix(INTSXP, (int)Gk, (int)N));
    SET_VECTOR_ELT(out, 1, Rf_allocMatrix(INTSXP, (int)Gk, (int)N));
    memcpy(INTEGER(VECTOR_ELT(out, 0)), o_r, (size_t)Gk * N * sizeof(int32_t));
    memcpy(INTEGER(VECTOR_ELT(out, 1)), o_n, (size_t)Gk * N * sizeof(int32_t));
    for (int i = 2; i < 5; i++) SET_VECTOR_ELT(out, i, Rf_allocVector(REALSXP, (R_xlen_t)Gk));
    memcpy(REAL(VECTOR_ELT(out, 2)), l, (size_t)Gk * sizeof(double));
    memcpy(REAL(VECTOR_ELT(out, 3)), nb, (size_t)Gk * sizeof(double));
    memcpy(REAL(VECTOR_ELT(out, 4)), lm, (size_t)Gk * sizeof(double));
    SET_VECTOR_ELT(out, 5, Rf_allocVector(INTSXP, (R_xlen_t)Gk));
    for (int64_t k = 0; k < Gk; k++) INTEGER(VECTOR_ELT(out, 5))[k] = keep[k] + 1;
    SET_VECTOR_ELT(out, 6, Rf_ScalarInteger((int)nhet));
    UNPROTECT(5);
    return out;
}

/* pm <- .Call(C_hb_retest_allele, r.maf[snps, ], n.sc[snps, ], l.maf[snps], n.bulk[snps], geneFactor, pe, mono):
 * the per-cell marginal of inst/bug/snpModel.bug instead of the JAGS run of calcAlleleCnvProb
 * (R/HoneyBADGER.R:907-1042).  geneFactor: integer codes per SNP (as.integer(factor(...))) or NULL. */
SEXP C_hb_retest_allele(SEXP r_rows, SEXP n_rows, SEXP l_maf, SEXP n_bulk, SEXP gene, SEXP pe, SEXP mono)
{
    r_rows = PROTECT(Rf_coerceVector(r_rows, INTSXP));
    n_rows = PROTECT(Rf_coerceVector(n_rows, INTSXP));
    l_maf = PROTECT(Rf_coerceVector(l_maf, REALSXP));
    n_bulk = PROTECT(Rf_coerceVector(n_bulk, REALSXP));
    SEXP gf = PROTECT(Rf_isNull(gene) ? R_NilValue : Rf_coerceVector(gene, INTSXP));
    const int64_t n = Rf_nrows(r_rows), N = Rf_ncols(r_rows);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 1, (int)N)); /* calcAlleleCnvProb returns a 1 x K matrix */
    int rc = hb_retest_allele_model_host(INTEGER(r_rows), INTEGER(n_rows), n, N, REAL(l_maf), REAL(n_bulk),
                                         Rf_isNull(gene) ? NULL : INTEGER(gf), Rf_asReal(pe), Rf_asReal(mono),
                                         REAL(out));
    UNPROTECT(6);
    hb_raise(rc);
    return out;
}

/* list(y, gamma, LL) <- .Call(C_hb_posteriors_norm, x, mean, sd, Pi, delta)
 * Viterbi states plus the state posteriors gamma [T x B x 3] = exp(logalpha + logbeta - LL) of
 * HiddenMarkov::forwardback and the log-likelihoods (north_star's posterior output; the reference never
 * computes it).  A pooled vector (one long chain) runs parallel along the gene axis. */
SEXP C_hb_posteriors_norm(SEXP x, SEXP mean, SEXP sd, SEXP Pi, SEXP delta)
{
    x = PROTECT(Rf_coerceVector(x, REALSXP));
    const int is_mat = Rf_isMatrix(x);
    const int64_t T = is_mat ? Rf_nrows(x) : XLENGTH(x), B = is_mat ? Rf_ncols(x) : 1;
    double pi_rm[9];
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) pi_rm[j * 3 + k] = REAL(Pi)[k * 3 + j];
    SEXP sdv = PROTECT(Rf_allocVector(REALSXP, B));
    for (int64_t b = 0; b < B; b++) REAL(sdv)[b] = REAL(sd)[XLENGTH(sd) == B ? b : 0];
    SEXP y = PROTECT(Rf_allocMatrix(INTSXP, (int)T, (int)B));
    SEXP g = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)T * B * 3)); /* [T x B x 3], dim set by the R wrapper */
    SEXP ll = PROTECT(Rf_allocVector(REALSXP, B));
    int rc = hb_hmm_norm_host(REAL(x), T, B, NULL, 1, REAL(mean), REAL(sdv), 0, pi_rm, REAL(delta),
                              HB_WANT_VITERBI | HB_WANT_GAMMA | HB_WANT_LL, INTEGER(y), REAL(g), REAL(ll));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, y);
    SET_VECTOR_ELT(out, 1, g);
    SET_VECTOR_ELT(out, 2, ll);
    UNPROTECT(6);
    hb_raise(rc);
    return out;
}

/* list(y, gamma, LL) <- .Call(C_hb_posteriors_binom, x, size, prob, Pi, delta): gamma [T x B x 2] */
SEXP C_hb_posteriors_binom(SEXP x, SEXP size, SEXP prob, SEXP Pi, SEXP delta)
{
    x = PROTECT(Rf_coerceVector(x, INTSXP));
    size = PROTECT(Rf_coerceVector(size, INTSXP));
    const int is_mat = Rf_isMatrix(x);
    const int64_t T = is_mat ? Rf_nrows(x) : XLENGTH(x), B = is_mat ? Rf_ncols(x) : 1;
    double pi_rm[4] = {REAL(Pi)[0], REAL(Pi)[2], REAL(Pi)[1], REAL(Pi)[3]};
    SEXP y = PROTECT(Rf_allocMatrix(INTSXP, (int)T, (int)B));
    SEXP g = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)T * B * 2));
    SEXP ll = PROTECT(Rf_allocVector(REALSXP, B));
    int rc = hb_hmm_binom_host(INTEGER(x), INTEGER(size), T, B, NULL, 1, REAL(prob), 0.0, pi_rm, REAL(delta),
                               HB_WANT_VITERBI | HB_WANT_GAMMA | HB_WANT_LL, INTEGER(y), REAL(g), REAL(ll));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, y);
    SET_VECTOR_ELT(out, 1, g);
    SET_VECTOR_ELT(out, 2, ll);
    UNPROTECT(6);
    hb_raise(rc);
    return out;
}

/* ------------------------------------------------------------------ device-resident handle
 * The matrices cross PCIe once; the RefClass keeps the external pointer in a field (`.dev`) for the whole
 * recursion (R/HoneyBADGER.R:1098-1102 keeps the matrices themselves there), and every level passes 0-based
 * index sets instead of copying `gexp.norm[, g1]` (:1304-1309).  The finalizer frees the device memory when
 * the R object is collected (or at exit). */
static void hb_handle_finalizer(SEXP ptr)
{
    hb_handle *h = (hb_handle *)R_ExternalPtrAddr(ptr);
    if (h) {
        hb_free_handle(h);
        R_ClearExternalPtr(ptr);
    }
}

static hb_handle *hb_handle_of(SEXP ptr)
{
    hb_handle *h = TYPEOF(ptr) == EXTPTRSXP ? (hb_handle *)R_ExternalPtrAddr(ptr) : NULL;
    if (!h) Rf_error("libhb_b200: invalid or released device handle");
    return h;
}

static SEXP hb_wrap_handle(hb_handle *h)
{
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, Rf_install("hb_handle"), R_NilValue));
    R_RegisterCFinalizerEx(ptr, hb_handle_finalizer, 1);
    UNPROTECT(1);
    return ptr;
}

/* idx: NULL (all) or a 1-based integer vector -> 0-based copy in R-managed memory */
static const int *hb_index0(SEXP idx, SEXP *keep, int64_t *n)
{
    if (Rf_isNull(idx)) {
        *keep = R_NilValue;
        *n = 0;
        return NULL;
    }
    SEXP src = PROTECT(Rf_coerceVector(idx, INTSXP));
    *n = XLENGTH(src);
    *keep = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)*n));
    for (int64_t i = 0; i < *n; i++) INTEGER(*keep)[i] = INTEGER(src)[i] - 1;
    UNPROTECT(2); /* callers PROTECT *keep again */
    return INTEGER(*keep);
}

/* dev <- .Call(C_hb_upload_gexp, gexp.norm)   (once, in setGexpMats / at init = TRUE) */
SEXP C_hb_upload_gexp(SEXP gexp_norm)
{
    gexp_norm = PROTECT(Rf_coerceVector(gexp_norm, REALSXP));
    hb_handle *h = NULL;
    int rc = hb_upload_gexp_host(REAL(gexp_norm), Rf_nrows(gexp_norm), Rf_ncols(gexp_norm), &h);
    UNPROTECT(1);
    hb_raise(rc);
    return hb_wrap_handle(h);
}

/* dev <- .Call(C_hb_upload_counts, r.maf, n.sc) */
SEXP C_hb_upload_counts(SEXP r_maf, SEXP n_sc)
{
    r_maf = PROTECT(Rf_coerceVector(r_maf, REALSXP));
    n_sc = PROTECT(Rf_coerceVector(n_sc, REALSXP));
    hb_handle *h = NULL;
    int rc = hb_upload_counts_host(REAL(r_maf), REAL(n_sc), Rf_nrows(r_maf), Rf_ncols(r_maf), &h);
    UNPROTECT(2);
    hb_raise(rc);
    return hb_wrap_handle(h);
}

/* .Call(C_hb_free_handle, dev): release now instead of waiting for the garbage collector */
SEXP C_hb_free_handle(SEXP ptr)
{
    hb_handle_finalizer(ptr);
    return R_NilValue;
}

/* bytes <- .Call(C_hb_handle_bytes, dev): what the upload moved host -> device */
SEXP C_hb_handle_bytes(SEXP ptr)
{
    int64_t b = 0;
    hb_raise(hb_handle_info(hb_handle_of(ptr), NULL, NULL, NULL, &b));
    return Rf_ScalarReal((double)b);
}

/* y <- .Call(C_hb_gexp_round_dev, dev, rows, cols, member, mean, Pi, delta): one lapply(heights, ...) round of
 * R/HoneyBADGER.R:1135-1161 on the resident matrix restricted to `rows` (genes) x `cols` (cells), 1-based or
 * NULL; member: logical / integer [length(cols) x K]; returns the [length(rows) x K] state matrix. */
SEXP C_hb_gexp_round_dev(SEXP ptr, SEXP rows, SEXP cols, SEXP member, SEXP mean, SEXP Pi, SEXP delta)
{
    hb_handle *h = hb_handle_of(ptr);
    int64_t G = 0, N = 0, nr = 0, nc = 0;
    hb_raise(hb_handle_info(h, NULL, &G, &N, NULL));
    SEXP kr, kc;
    const int *r0 = hb_index0(rows, &kr, &nr);
    PROTECT(kr);
    const int *c0 = hb_index0(cols, &kc, &nc);
    PROTECT(kc);
    const int64_t G2 = r0 ? nr : G, N2 = c0 ? nc : N;
    member = PROTECT(Rf_coerceVector(member, INTSXP));
    const int K = Rf_ncols(member);
    SEXP mem8 = PROTECT(Rf_allocVector(LGLSXP, (R_xlen_t)((K * N2 + 3) / 4 + 1))); /* scratch bytes */
    unsigned char *m8 = (unsigned char *)LOGICAL(mem8);
    for (int64_t i = 0; i < (int64_t)K * N2; i++) m8[i] = INTEGER(member)[i] != 0; /* [N2 x K] column-major == [K][N2] */
    double pi_rm[9];
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) pi_rm[j * 3 + k] = REAL(Pi)[k * 3 + j];
    SEXP y = PROTECT(Rf_allocMatrix(INTSXP, (int)G2, K)); /* [K][G2] row-major == [G2 x K] column-major */
    int rc = hb_gexp_round_handle(h, r0, nr, c0, nc, m8, K, REAL(mean), pi_rm, REAL(delta), INTEGER(y), NULL, NULL);
    UNPROTECT(5);
    hb_raise(rc);
    return y;
}

/* res <- .Call(C_hb_allele_round_dev, dev, rows, cols, member, prob, Pi, delta): R/HoneyBADGER.R:1395-1417;
 * list(y, mafl, sizel), each [length(rows) x K]. */
SEXP C_hb_allele_round_dev(SEXP ptr, SEXP rows, SEXP cols, SEXP member, SEXP prob, SEXP Pi, SEXP delta)
{
    hb_handle *h = hb_handle_of(ptr);
    int64_t G = 0, N = 0, nr = 0, nc = 0;
    hb_raise(hb_handle_info(h, NULL, &G, &N, NULL));
    SEXP kr, kc;
    const int *r0 = hb_index0(rows, &kr, &nr);
    PROTECT(kr);
    const int *c0 = hb_index0(cols, &kc, &nc);
    PROTECT(kc);
    const int64_t G2 = r0 ? nr : G, N2 = c0 ? nc : N;
    member = PROTECT(Rf_coerceVector(member, INTSXP));
    const int K = Rf_ncols(member);
    SEXP mem8 = PROTECT(Rf_allocVector(LGLSXP, (R_xlen_t)((K * N2 + 3) / 4 + 1)));
    unsigned char *m8 = (unsigned char *)LOGICAL(mem8);
    for (int64_t i = 0; i < (int64_t)K * N2; i++) m8[i] = INTEGER(member)[i] != 0;
    double pi_rm[4] = {REAL(Pi)[0], REAL(Pi)[2], REAL(Pi)[1], REAL(Pi)[3]};
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    for (int i = 0; i < 3; i++) SET_VECTOR_ELT(out, i, Rf_allocMatrix(INTSXP, (int)G2, K));
    int rc = hb_allele_round_handle(h, r0, nr, c0, nc, m8, K, REAL(prob), pi_rm, REAL(delta),
                                    INTEGER(VECTOR_ELT(out, 0)), INTEGER(VECTOR_ELT(out, 1)),
                                    INTEGER(VECTOR_ELT(out, 2)));
    UNPROTECT(5);
    hb_raise(rc);
    return out;
}

/* ct <- .Call(C_hb_group_dev, dev, rows, cols, window, heights): cutree labels [length(cols) x length(heights)] */
SEXP C_hb_group_dev(SEXP ptr, SEXP rows, SEXP cols, SEXP window, SEXP heights)
{
    hb_handle *h = hb_handle_of(ptr);
    int64_t N = 0, nr = 0, nc = 0;
    hb_raise(hb_handle_info(h, NULL, NULL, &N, NULL));
    SEXP kr, kc;
    const int *r0 = hb_index0(rows, &kr, &nr);
    PROTECT(kr);
    const int *c0 = hb_index0(cols, &kc, &nc);
    PROTECT(kc);
    heights = PROTECT(Rf_coerceVector(heights, INTSXP));
    const int nk = (int)XLENGTH(heights);
    SEXP lab = PROTECT(Rf_allocMatrix(INTSXP, (int)(c0 ? nc : N), nk));
    int rc = hb_group_handle(h, r0, nr, c0, nc, Rf_asInteger(window), INTEGER(heights), nk, INTEGER(lab));
    UNPROTECT(4);
    hb_raise(rc);
    return lab;
}

/* p <- .Call(C_hb_retest_gexp_dev, dev, rows, cols, m, sigma0): list(amp, del, mu0) over `cols` for region `rows` */
SEXP C_hb_retest_gexp_dev(SEXP ptr, SEXP rows, SEXP cols, SEXP mag0, SEXP sigma0)
{
    hb_handle *h = hb_handle_of(ptr);
    int64_t N = 0, nr = 0, nc = 0;
    hb_raise(hb_handle_info(h, NULL, NULL, &N, NULL));
    SEXP kr, kc;
    const int *r0 = hb_index0(rows, &kr, &nr);
    PROTECT(kr);
    const int *c0 = hb_index0(cols, &kc, &nc);
    PROTECT(kc);
    const int64_t N2 = c0 ? nc : N;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    for (int i = 0; i < 3; i++) SET_VECTOR_ELT(out, i, Rf_allocVector(REALSXP, (R_xlen_t)N2));
    int rc = hb_retest_gexp_handle(h, r0, nr, c0, nc, Rf_asReal(mag0), Rf_isNull(sigma0) ? NULL : REAL(sigma0),
                                   REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)));
    UNPROTECT(3);
    hb_raise(rc);
    return out;
}

/* pm <- .Call(C_hb_retest_allele_dev, dev, rows, cols, l.maf, n.bulk, gene, pe, mono): the snpModel marginal of
 * R/HoneyBADGER.R:1486 on the resident counts; rows already grouped by gene, `gene` = integer factor codes of the
 * rows (NULL: one gene per SNP).  Returns the 1 x length(cols) matrix calcAlleleCnvProb returns. */
SEXP C_hb_retest_allele_dev(SEXP ptr, SEXP rows, SEXP cols, SEXP l_maf, SEXP n_bulk, SEXP gene, SEXP pe, SEXP mono)
{
    hb_handle *h = hb_handle_of(ptr);
    int64_t G = 0, N = 0, nr = 0, nc = 0;
    hb_raise(hb_handle_info(h, NULL, &G, &N, NULL));
    SEXP kr, kc;
    const int *r0 = hb_index0(rows, &kr, &nr);
    PROTECT(kr);
    const int *c0 = hb_index0(cols, &kc, &nc);
    PROTECT(kc);
    const int64_t G2 = r0 ? nr : G, N2 = c0 ? nc : N;
    l_maf = PROTECT(Rf_coerceVector(l_maf, REALSXP));
    n_bulk = PROTECT(Rf_coerceVector(n_bulk, REALSXP));
    unsigned char *ge = NULL;
    if (!Rf_isNull(gene)) { /* last SNP of every run of equal codes */
        gene = Rf_coerceVector(gene, INTSXP);
        ge = (unsigned char *)R_alloc((size_t)G2, 1);
        for (int64_t j = 0; j < G2; j++) ge[j] = (j + 1 == G2) || INTEGER(gene)[j + 1] != INTEGER(gene)[j];
    }
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 1, (int)N2));
    int rc = hb_retest_allele_handle(h, r0, nr, c0, nc, REAL(l_maf), REAL(n_bulk), ge, Rf_asReal(pe), Rf_asReal(mono),
                                     REAL(out));
    UNPROTECT(5);
    hb_raise(rc);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_hb_viterbi_norm", (DL_FUNC)&C_hb_viterbi_norm, 5},
    {"C_hb_viterbi_binom", (DL_FUNC)&C_hb_viterbi_binom, 5},
    {"C_hb_posteriors_norm", (DL_FUNC)&C_hb_posteriors_norm, 5},
    {"C_hb_posteriors_binom", (DL_FUNC)&C_hb_posteriors_binom, 5},
    {"C_hb_gexp_pooled_viterbi", (DL_FUNC)&C_hb_gexp_pooled_viterbi, 5},
    {"C_hb_allele_pooled_viterbi", (DL_FUNC)&C_hb_allele_pooled_viterbi, 6},
    {"C_hb_group_gexp", (DL_FUNC)&C_hb_group_gexp, 3},
    {"C_hb_group_allele", (DL_FUNC)&C_hb_group_allele, 4},
    {"C_hb_retest_gexp", (DL_FUNC)&C_hb_retest_gexp, 3},
    {"C_hb_retest_allele", (DL_FUNC)&C_hb_retest_allele, 7},
    {"C_hb_set_gexp_mats", (DL_FUNC)&C_hb_set_gexp_mats, 7},
    {"C_hb_set_allele_mats", (DL_FUNC)&C_hb_set_allele_mats, 7},
    {"C_hb_upload_gexp", (DL_FUNC)&C_hb_upload_gexp, 1},
    {"C_hb_upload_counts", (DL_FUNC)&C_hb_upload_counts, 2},
    {"C_hb_free_handle", (DL_FUNC)&C_hb_free_handle, 1},
    {"C_hb_handle_bytes", (DL_FUNC)&C_hb_handle_bytes, 1},
    {"C_hb_gexp_round_dev", (DL_FUNC)&C_hb_gexp_round_dev, 7},
    {"C_hb_allele_round_dev", (DL_FUNC)&C_hb_allele_round_dev, 7},
    {"C_hb_group_dev", (DL_FUNC)&C_hb_group_dev, 5},
    {"C_hb_retest_gexp_dev", (DL_FUNC)&C_hb_retest_gexp_dev, 5},
    {"C_hb_retest_allele_dev", (DL_FUNC)&C_hb_retest_allele_dev, 8},
    {NULL, NULL, 0}};

void R_init_hbb200(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
