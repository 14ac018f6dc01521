/* Minimal declarations of the R C API used by r-package/src/npdrb200_shim.c -- enough for a -fsyntax-only check in a
 * container without R (tests/test_host.py::test_r_shim_compiles_against_api_stub).  Not R's headers. */
#include <stddef.h>
typedef struct SEXPREC *SEXP; typedef ptrdiff_t R_xlen_t; typedef void *(*DL_FUNC)(void);
extern SEXP R_DimSymbol, R_NilValue;
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
SEXP Rf_getAttrib(SEXP, SEXP); int *INTEGER(SEXP); double *REAL(SEXP); SEXP Rf_allocMatrix(int, int, int); SEXP Rf_allocVector(int, R_xlen_t);
SEXP Rf_protect(SEXP); void Rf_unprotect(int); void Rf_error(const char *, ...); int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_asLogical(SEXP);
int Rf_isNull(SEXP); SEXP Rf_mkNamed(int, const char **); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); SEXP Rf_ScalarReal(double); R_xlen_t XLENGTH(SEXP);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
typedef void (*R_CFinalizer_t)(SEXP); typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
SEXP R_MakeExternalPtr(void *, SEXP, SEXP); void *R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean); char *R_alloc(size_t, int); SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
