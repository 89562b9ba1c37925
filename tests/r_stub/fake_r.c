/* TEST-ONLY miniature of the part of R's C API that r/hb_shim.c uses, so that the shim — the file a
 * HoneyBADGER maintainer adds to the package — can be LINKED AND RUN against libhb_b200.so without R
 * (tests/test_r_shim_runs.py).  It is not R: no garbage collector (objects live until fr_reset), PROTECT is a
 * counter whose balance the tests check, Rf_error longjmps back to fr_call like R's error unwinding does. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
    SEXPTYPE type;
    R_xlen_t len;
    int nrow, ncol; /* -1: not a matrix */
    void *data;
    R_CFinalizer_t fin;
};

static struct SEXPREC nil_obj = {NILSXP, 0, -1, -1, NULL, NULL};
SEXP R_NilValue = &nil_obj;

static SEXP *objs = NULL;
static size_t nobjs = 0, cap = 0;
static void **ralloc = NULL;
static size_t nralloc = 0, capr = 0;
static int protect_depth = 0, depth_at_error = -1, depth_at_entry = 0;
static char errmsg[1024];
static jmp_buf *cur_jmp = NULL;
static const R_CallMethodDef *table = NULL;

static size_t elsize(SEXPTYPE t)
{
    switch (t) {
    case REALSXP: return sizeof(double);
    case INTSXP:
    case LGLSXP: return sizeof(int);
    case VECSXP: return sizeof(SEXP);
    default: return 1;
    }
}

static SEXP newobj(SEXPTYPE t, R_xlen_t n)
{
    SEXP s = (SEXP)calloc(1, sizeof *s);
    s->type = t;
    s->len = n;
    s->nrow = s->ncol = -1;
    s->data = calloc((size_t)(n > 0 ? n : 1), elsize(t));
    if (t == VECSXP)
        for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
    if (nobjs == cap) {
        cap = cap ? 2 * cap : 64;
        objs = (SEXP *)realloc(objs, cap * sizeof *objs);
    }
    objs[nobjs++] = s;
    return s;
}

/* ------------------------------------------------------------------ the R API subset */
SEXP Rf_protect(SEXP s) { protect_depth++; return s; }
void Rf_unprotect(int n) { protect_depth -= n; }
double *REAL(SEXP s) { return (double *)s->data; }
int *INTEGER(SEXP s) { return (int *)s->data; }
int *LOGICAL(SEXP s) { return (int *)s->data; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP *)s->data)[i] = v; return v; }
SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) { return newobj(t, n); }
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc)
{
    SEXP s = newobj(t, (R_xlen_t)nr * nc);
    s->nrow = nr;
    s->ncol = nc;
    return s;
}
int Rf_nrows(SEXP s) { return s->nrow >= 0 ? s->nrow : (int)s->len; }
int Rf_ncols(SEXP s) { return s->ncol >= 0 ? s->ncol : 1; }
int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
int Rf_isMatrix(SEXP s) { return s->nrow >= 0; }
int TYPEOF(SEXP s) { return (int)s->type; }
double Rf_asReal(SEXP s) { return s->type == REALSXP ? REAL(s)[0] : (double)INTEGER(s)[0]; }
int Rf_asInteger(SEXP s) { return s->type == REALSXP ? (int)REAL(s)[0] : INTEGER(s)[0]; }
int Rf_asLogical(SEXP s) { return Rf_asInteger(s) != 0; }
SEXP Rf_ScalarInteger(int v) { SEXP s = newobj(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = newobj(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_install(const char *name) { (void)name; return R_NilValue; }
SEXP Rf_coerceVector(SEXP s, SEXPTYPE t)
{
    if (s->type == t || Rf_isNull(s)) return s;
    SEXP r = newobj(t, s->len);
    r->nrow = s->nrow;
    r->ncol = s->ncol;
    for (R_xlen_t i = 0; i < s->len; i++) {
        double v = s->type == REALSXP ? REAL(s)[i] : (double)INTEGER(s)[i];
        if (t == REALSXP) REAL(r)[i] = v;
        else INTEGER(r)[i] = (int)v; /* R truncates; the shim only coerces integer-valued doubles */
    }
    return r;
}
char *R_alloc(size_t n, int size)
{
    if (nralloc == capr) {
        capr = capr ? 2 * capr : 16;
        ralloc = (void **)realloc(ralloc, capr * sizeof *ralloc);
    }
    return (char *)(ralloc[nralloc++] = calloc(n ? n : 1, (size_t)size));
}
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, sizeof errmsg, fmt, ap);
    va_end(ap);
    depth_at_error = protect_depth;
    if (!cur_jmp) {
        fprintf(stderr, "fake R: error outside fr_call: %s\n", errmsg);
        abort();
    }
    longjmp(*cur_jmp, 1);
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag, (void)prot;
    SEXP s = newobj(EXTPTRSXP, 0);
    free(s->data);
    s->data = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, int onexit) { (void)onexit; s->fin = fun; }
int R_registerRoutines(DllInfo *d, const R_CMethodDef *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)d, (void)c, (void)f, (void)e;
    table = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *d, Rboolean v) { (void)d; return v; }

/* ------------------------------------------------------------------ harness (called through ctypes) */
void R_init_hbb200(DllInfo *dll);

int fr_init(void) { R_init_hbb200(NULL); return table != NULL; }
int fr_arity(const char *name)
{
    for (const R_CallMethodDef *m = table; m && m->name; m++)
        if (!strcmp(m->name, name)) return m->numArgs;
    return -1;
}
SEXP fr_nil(void) { return R_NilValue; }
SEXP fr_real(R_xlen_t n, const double *v) { SEXP s = newobj(REALSXP, n); if (v) memcpy(s->data, v, (size_t)n * 8); return s; }
SEXP fr_int(R_xlen_t n, const int *v) { SEXP s = newobj(INTSXP, n); if (v) memcpy(s->data, v, (size_t)n * 4); return s; }
SEXP fr_lgl(R_xlen_t n, const int *v) { SEXP s = newobj(LGLSXP, n); if (v) memcpy(s->data, v, (size_t)n * 4); return s; }
void fr_set_dim(SEXP s, int nr, int nc) { s->nrow = nr; s->ncol = nc; }
int fr_type(SEXP s) { return (int)s->type; }
long long fr_len(SEXP s) { return (long long)s->len; }
int fr_nrow(SEXP s) { return s->nrow; }
int fr_ncol(SEXP s) { return s->ncol; }
void *fr_data(SEXP s) { return s->data; }
SEXP fr_elt(SEXP s, long long i) { return VECTOR_ELT(s, (R_xlen_t)i); }
const char *fr_error_message(void) { return errmsg; }
int fr_protect_depth(void) { return protect_depth; }
int fr_depth_at_error(void) { return depth_at_error - depth_at_entry; }
void fr_finalize(SEXP s) { if (s->type == EXTPTRSXP && s->fin) s->fin(s); }

/* .Call(name, args...): NULL on Rf_error (message in fr_error_message); the PROTECT stack is unwound to the
 * entry depth on error, as R does */
SEXP fr_call(const char *name, int nargs, SEXP *a)
{
    const R_CallMethodDef *volatile m = table;
    for (; m && m->name; m++)
        if (!strcmp(m->name, name)) break;
    if (!m || !m->name || m->numArgs != nargs) {
        snprintf(errmsg, sizeof errmsg, "no routine %s with %d arguments", name, nargs);
        return NULL;
    }
    jmp_buf jb;
    cur_jmp = &jb;
    depth_at_entry = protect_depth;
    depth_at_error = -1;
    errmsg[0] = 0;
    SEXP volatile r = NULL;
    if (setjmp(jb) == 0) {
        typedef SEXP (*f1)(SEXP);
        typedef SEXP (*f2)(SEXP, SEXP);
        typedef SEXP (*f3)(SEXP, SEXP, SEXP);
        typedef SEXP (*f4)(SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f5)(SEXP, SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
        switch (nargs) {
        case 1: r = ((f1)m->fun)(a[0]); break;
        case 2: r = ((f2)m->fun)(a[0], a[1]); break;
        case 3: r = ((f3)m->fun)(a[0], a[1], a[2]); break;
        case 4: r = ((f4)m->fun)(a[0], a[1], a[2], a[3]); break;
        case 5: r = ((f5)m->fun)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: r = ((f6)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: r = ((f7)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: r = ((f8)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        default: snprintf(errmsg, sizeof errmsg, "arity %d not wired", nargs); break;
        }
    } else {
        protect_depth = depth_at_entry;
        r = NULL;
    }
    cur_jmp = NULL;
    for (size_t i = 0; i < nralloc; i++) free(ralloc[i]);
    nralloc = 0;
    return r;
}

/* free every object made since the last reset (external pointers are finalized first) */
void fr_reset(void)
{
    for (size_t i = 0; i < nobjs; i++)
        if (objs[i]->type == EXTPTRSXP && objs[i]->fin && objs[i]->data) objs[i]->fin(objs[i]);
    for (size_t i = 0; i < nobjs; i++) {
        if (objs[i]->type != EXTPTRSXP) free(objs[i]->data);
        free(objs[i]);
    }
    nobjs = 0;
    protect_depth = 0;
}
