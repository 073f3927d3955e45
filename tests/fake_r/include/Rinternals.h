/* TEST INFRASTRUCTURE: a stand-in for the slice of R's C API that gpirt_rshim.c uses, so the shim can be compiled and
 * executed in an image without R (tests/fake_r/fake_r.c implements it; tests/test_rshim.py drives it). */
#ifndef FAKE_RINTERNALS_H
#define FAKE_RINTERNALS_H
#include <stddef.h>
#include <limits.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define CHARSXP 9
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define NA_INTEGER INT_MIN
typedef struct fake_sexp {
    SEXPTYPE type;
    R_xlen_t length;
    void* data;                 /* double[], int[], SEXP[] or char[] */
    int dims[3];
    int ndim;
    struct fake_sexp* names;
} *SEXP;
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))
#define TYPEOF(x) ((x)->type)
#define XLENGTH(x) ((x)->length)
#define LENGTH(x) ((int)(x)->length)
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_alloc3DArray(SEXPTYPE, int, int, int);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_mkChar(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
int Rf_isMatrix(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
const char* CHAR(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
void Rf_onintr(void);
#ifdef __cplusplus
}
#endif
#endif
