/* Fake R runtime behind tests/r_stub/Rinternals.h + the harness the Python tests drive through ctypes - TEST INFRASTRUCTURE. */
#include "Rinternals.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, 0, 0, NULL, NULL, NULL};
static struct SEXPREC names_rec = {NILSXP, 0, NULL, NULL, 0, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec;
SEXP R_NamesSymbol = &names_rec;
static jmp_buf g_top;
static int g_top_set = 0;
static char g_err[1024];
static int g_protect_depth = 0;

static size_t elt_size(int type) {
  switch (type) {
    case REALSXP: return sizeof(double);
    case INTSXP: return sizeof(int);
    case RAWSXP: case CHARSXP: return 1;
    case STRSXP: case VECSXP: return sizeof(SEXP);
    default: return 0;
  }
}
static SEXP new_rec(int type, R_xlen_t n) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type; s->len = n; s->names = R_NilValue; s->prot = R_NilValue;
  const size_t es = elt_size(type);
  if (es) s->data = calloc((size_t)(n > 0 ? n : 1) + (type == CHARSXP), es);
  if (type == STRSXP || type == VECSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
double* REAL(SEXP x) { if (x->type != REALSXP) Rf_error("REAL() on type %d", x->type); return (double*)x->data; }
int* INTEGER(SEXP x) { if (x->type != INTSXP) Rf_error("INTEGER() on type %d", x->type); return (int*)x->data; }
Rbyte* RAW(SEXP x) { if (x->type != RAWSXP) Rf_error("RAW() on type %d", x->type); return (Rbyte*)x->data; }
const char* CHAR(SEXP x) { return (const char*)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
int Rf_asInteger(SEXP x) { return x->type == INTSXP ? INTEGER(x)[0] : (int)REAL(x)[0]; }
double Rf_asReal(SEXP x) { return x->type == INTSXP ? (double)INTEGER(x)[0] : REAL(x)[0]; }
int Rf_nrows(SEXP x) { return x->nrow ? x->nrow : (int)x->len; }
int Rf_ncols(SEXP x) { return x->ncol ? x->ncol : 1; }
int Rf_length(SEXP x) { return (int)x->len; }
R_xlen_t Rf_xlength(SEXP x) { return x->len; }
SEXP Rf_allocVector(unsigned type, R_xlen_t n) { return new_rec((int)type, n); }
SEXP Rf_allocMatrix(unsigned type, int nrow, int ncol) {
  SEXP s = new_rec((int)type, (R_xlen_t)nrow * ncol);
  s->nrow = nrow; s->ncol = ncol;
  return s;
}
SEXP Rf_getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : R_NilValue; }
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag;
  SEXP s = new_rec(EXTPTRSXP, 0);
  s->ptr = p; s->prot = prot;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->ptr : NULL; }
void R_ClearExternalPtr(SEXP s) { s->ptr = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }
SEXP Rf_protect(SEXP x) { ++g_protect_depth; return x; }
void Rf_unprotect(int n) { g_protect_depth -= n; }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  if (!g_top_set) { fprintf(stderr, "Rf_error outside stub_call: %s\n", g_err); abort(); }
  longjmp(g_top, 1);
}

/* ---- harness (called from Python) ------------------------------------------------------------------------------------ */
SEXP stub_nil(void) { return R_NilValue; }
SEXP stub_real(const double* v, R_xlen_t n, int nrow, int ncol) {
  SEXP s = new_rec(REALSXP, n);
  if (n) memcpy(s->data, v, (size_t)n * sizeof(double));
  s->nrow = nrow; s->ncol = ncol;
  return s;
}
SEXP stub_int(const int* v, R_xlen_t n) { SEXP s = new_rec(INTSXP, n); if (n) memcpy(s->data, v, (size_t)n * sizeof(int)); return s; }
SEXP stub_raw(const void* v, R_xlen_t n) { SEXP s = new_rec(RAWSXP, n); if (n) memcpy(s->data, v, (size_t)n); return s; }
static SEXP mkchar(const char* c) { SEXP s = new_rec(CHARSXP, (R_xlen_t)strlen(c)); strcpy((char*)s->data, c); return s; }
SEXP stub_string(const char* c) { SEXP s = new_rec(STRSXP, 1); ((SEXP*)s->data)[0] = mkchar(c); return s; }
SEXP stub_strings(const char** v, int n) { SEXP s = new_rec(STRSXP, n); for (int i = 0; i < n; ++i) ((SEXP*)s->data)[i] = mkchar(v[i]); return s; }
SEXP stub_list(int n) { SEXP s = new_rec(VECSXP, n); s->names = new_rec(STRSXP, n); return s; }
void stub_list_set(SEXP l, int i, const char* name, SEXP v) { ((SEXP*)l->data)[i] = v; ((SEXP*)l->names->data)[i] = mkchar(name); }
int stub_type(SEXP s) { return s->type; }
R_xlen_t stub_len(SEXP s) { return s->len; }
void* stub_data(SEXP s) { return s->data; }
int stub_nrow(SEXP s) { return s->nrow; }
int stub_ncol(SEXP s) { return s->ncol; }
SEXP stub_prot(SEXP s) { return s->prot; }
const char* stub_last_error(void) { return g_err; }
int stub_protect_depth(void) { return g_protect_depth; }
/* what the GC does to an unreachable external pointer: run its finalizer once */
void stub_finalize(SEXP s) { if (s->type == EXTPTRSXP && s->fin) { R_CFinalizer_t f = s->fin; s->fin = NULL; f(s); } }

typedef SEXP (*fn1)(SEXP); typedef SEXP (*fn2)(SEXP, SEXP); typedef SEXP (*fn3)(SEXP, SEXP, SEXP);
typedef SEXP (*fn4)(SEXP, SEXP, SEXP, SEXP); typedef SEXP (*fn6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn0)(void); typedef SEXP (*fn5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
/* .Call(fn, args...): returns the result, or NULL after an R error (message in stub_last_error) */
SEXP stub_call(void* fn, int nargs, SEXP* a) {
  g_err[0] = 0;
  g_protect_depth = 0;
  if (setjmp(g_top)) { g_top_set = 0; return NULL; }
  g_top_set = 1;
  SEXP r = NULL;
  switch (nargs) {
    case 0: r = ((fn0)fn)(); break;
    case 1: r = ((fn1)fn)(a[0]); break;
    case 2: r = ((fn2)fn)(a[0], a[1]); break;
    case 3: r = ((fn3)fn)(a[0], a[1], a[2]); break;
    case 4: r = ((fn4)fn)(a[0], a[1], a[2], a[3]); break;
    case 5: r = ((fn5)fn)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: r = ((fn6)fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 10: r = ((fn10)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
    case 9: r = ((fn9)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
    default: snprintf(g_err, sizeof g_err, "stub_call: unsupported arity %d", nargs); r = NULL;
  }
  g_top_set = 0;
  return r;
}
