/* Minimal stand-in for the parts of R's C API that r_glue/ uses -- declarations
 * only, for `g++ -fsyntax-only` (tests/test_r_glue.py).  TEST INFRASTRUCTURE. */
#pragma once
#include <cstddef>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
extern SEXP R_NilValue;
enum { NILSXP = 0, INTSXP = 13, REALSXP = 14, RAWSXP = 24 };
int TYPEOF(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
Rbyte* RAW(SEXP);
int Rf_length(SEXP);
R_xlen_t Rf_xlength(SEXP);
void* R_ExternalPtrAddr(SEXP);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
