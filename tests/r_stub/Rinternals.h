/* stub of Rinternals.h: just the declarations r_shim.c uses, with R's documented signatures */
#ifndef RINTERNALS_STUB_H
#define RINTERNALS_STUB_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define FALSE 0
#define TRUE 1
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NamesSymbol;
SEXP Rf_getAttrib(SEXP, SEXP);
R_xlen_t Rf_xlength(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *CHAR(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isMatrix(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
void Rf_warning(const char *, ...);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkNamed(unsigned int, const char **);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
Rboolean R_ToplevelExec(void (*fun)(void *), void *data);
void R_CheckUserInterrupt(void);
#endif
