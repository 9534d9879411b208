/* Declaration-only stand-in for <Rinternals.h> (see R.h in this directory). */
#ifndef STUB_RINTERNALS_H
#define STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef struct { double r; double i; } Rcomplex;
typedef unsigned int SEXPTYPE;
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define CPLXSXP 15
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_mkChar(const char*);
int Rf_isComplex(SEXP);
int Rf_isNull(SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
Rcomplex* COMPLEX(SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define allocMatrix Rf_allocMatrix
#define allocVector Rf_allocVector
#define getAttrib Rf_getAttrib
#define setAttrib Rf_setAttrib
#define mkChar Rf_mkChar
#define isComplex Rf_isComplex
#define isNull Rf_isNull
#define asReal Rf_asReal
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#endif
