/* Minimal DECLARATIONS of the R C API used by r-package/src/r_shim.c, so that the shim can be
 * syntax- and type-checked where R is not installed (tests/test_host_cpu.py).  Not R's headers:
 * nothing here is linked or executed. */
#ifndef NRB_R_STUBS_H
#define NRB_R_STUBS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef void (*R_CFinalizer_t)(SEXP);
extern SEXP R_NilValue;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
int *INTEGER(SEXP);
double *REAL(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_ScalarLogical(int);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
char *R_alloc(size_t, int);
int R_IsNA(double);
#define ISNA(x) R_IsNA(x)
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void R_CheckUserInterrupt(void);
#endif
