/* Stub of the slice of R's C API that r/mb_shim.c uses - TEST INFRASTRUCTURE (no R in the build image).
 * SEXPs are plain heap records; Rf_error longjmps back into stub_call() like R's error longjmps to the top level. */
#ifndef MB_STUB_RINTERNALS_H
#define MB_STUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef struct SEXPREC* SEXP;
typedef void (*R_CFinalizer_t)(SEXP);
enum { NILSXP = 0, CHARSXP = 9, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, RAWSXP = 24 };
struct SEXPREC {
  int type;
  R_xlen_t len;
  void* data;              /* double / int / Rbyte / char / SEXP array */
  SEXP names;              /* STRSXP or R_NilValue */
  int nrow, ncol;          /* dim attribute (0 = none) */
  void* ptr; SEXP prot; R_CFinalizer_t fin;   /* external pointers */
};
extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;
double* REAL(SEXP x);
int* INTEGER(SEXP x);
Rbyte* RAW(SEXP x);
const char* CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_length(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
SEXP Rf_allocVector(unsigned type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned type, int nrow, int ncol);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
#ifdef __cplusplus
}
#endif
#endif
