/* TEST-ONLY stand-in for <R_ext/Rdynload.h> (see ../R.h). */
#ifndef HB_RDYNLOAD_STUB_H
#define HB_RDYNLOAD_STUB_H
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; void *types; } R_CMethodDef;
typedef struct _DllInfo DllInfo;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
int R_registerRoutines(DllInfo *, const R_CMethodDef *, const R_CallMethodDef *, const void *, const void *);
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean);
#endif
