/* TEST-ONLY stand-in for R's <R.h>: just enough declarations for `gcc -fsyntax-only r/hb_shim.c`
 * (tests/test_host_logic.py).  R is not installed in this image; this checks that the shim's calls match
 * include/hb_b200.h and that its use of the R API is well-formed C, nothing more. */
#ifndef HB_R_STUB_H
#define HB_R_STUB_H
#include <stddef.h>
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
#endif
