/* TEST-ONLY stand-in for <Rinternals.h> (see R.h in this directory). */
#ifndef HB_RINTERNALS_STUB_H
#define HB_RINTERNALS_STUB_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_isNull(SEXP);
int Rf_isMatrix(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
int Rf_asInteger(SEXP);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
#define EXTPTRSXP 22
typedef void (*R_CFinalizer_t)(SEXP);
typedef enum { FALSE_ = 0, TRUE_ } Rboolean_stub;
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, int onexit);
SEXP Rf_install(const char *);
int TYPEOF(SEXP);
#endif
