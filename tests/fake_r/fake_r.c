/* TEST INFRASTRUCTURE: minimal R runtime stand-in + a harness that loads the shim the way R would
 * (R_init_gpirt -> registered .Call routine with 7 arguments) and calls it on a small problem read from a file.
 *   fake_r_harness <in.bin> <out.bin> [thin]      (thin > 0: the extended routine _gpirt_gpirtMCMC_b200, arity 10)
 * in.bin : int32 n, m, S, B; uint64 rng_state; double y[n*m], theta[n], pm[2m], psd[2m], pstep[2m]
 * out.bin: uint64 seed_used; theta (S+1)*n, beta 2*m*(S+1), f n*m*(S+1), IRFs 1001*m   (doubles)
 * exit code 0 ok, 3 = Rf_error was raised (message on stderr), 2 = registration problem */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Random.h>
#include <R_ext/Utils.h>

#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static struct fake_sexp nil_obj = {NILSXP, 0, NULL, {0, 0, 0}, 0, NULL}, names_sym, dim_sym;
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
static jmp_buf error_jmp;
static uint64_t rng_state = 0x853c49e6748fea9bULL;
static int rng_open = 0, quiet = 1;

static size_t elt_size(SEXPTYPE t) { return t == REALSXP ? sizeof(double) : t == INTSXP ? sizeof(int) : t == CHARSXP ? 1 : sizeof(SEXP); }
SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    s->type = t; s->length = n; s->data = calloc((size_t)(n > 0 ? n : 1), elt_size(t)); s->ndim = 1; s->dims[0] = (int)n;
    return s;
}
SEXP Rf_allocMatrix(SEXPTYPE t, int r, int c) { SEXP s = Rf_allocVector(t, (R_xlen_t)r * c); s->ndim = 2; s->dims[0] = r; s->dims[1] = c; return s; }
SEXP Rf_alloc3DArray(SEXPTYPE t, int a, int b, int c) { SEXP s = Rf_allocVector(t, (R_xlen_t)a * b * c); s->ndim = 3; s->dims[0] = a; s->dims[1] = b; s->dims[2] = c; return s; }
SEXP Rf_coerceVector(SEXP x, SEXPTYPE t) {
    if (x->type == t) return x;
    SEXP s = Rf_allocVector(t, x->length);
    memcpy(s->dims, x->dims, sizeof(s->dims)); s->ndim = x->ndim;
    for (R_xlen_t i = 0; i < x->length; ++i) {
        if (x->type == INTSXP && t == REALSXP) ((double*)s->data)[i] = ((int*)x->data)[i] == NA_INTEGER ? (0.0 / 0.0) : ((int*)x->data)[i];
        else if (x->type == REALSXP && t == INTSXP) ((int*)s->data)[i] = (int)((double*)x->data)[i];
    }
    return s;
}
SEXP Rf_mkChar(const char* c) { SEXP s = Rf_allocVector(CHARSXP, (R_xlen_t)strlen(c) + 1); strcpy((char*)s->data, c); return s; }
SEXP Rf_setAttrib(SEXP x, SEXP sym, SEXP v) { if (sym == R_NamesSymbol) x->names = v; return v; }
int Rf_isMatrix(SEXP x) { return x->ndim == 2; }
int Rf_nrows(SEXP x) { return x->dims[0]; }
int Rf_ncols(SEXP x) { return x->ndim >= 2 ? x->dims[1] : 1; }
int Rf_asInteger(SEXP x) { return x->type == INTSXP ? ((int*)x->data)[0] : (int)((double*)x->data)[0]; }
double* REAL(SEXP x) { return (double*)x->data; }
int* INTEGER(SEXP x) { return (int*)x->data; }
void SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
const char* CHAR(SEXP x) { return (const char*)x->data; }
void Rf_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); fprintf(stderr, "Error: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap);
    longjmp(error_jmp, 1);
}
void Rf_warning(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); fprintf(stderr, "Warning: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap);
}
void Rf_onintr(void) { fprintf(stderr, "Interrupted\n"); longjmp(error_jmp, 2); }
void Rprintf(const char* fmt, ...) { if (quiet) return; va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); }
void GetRNGstate(void) { rng_open = 1; }
void PutRNGstate(void) { rng_open = 0; }
double unif_rand(void) {   /* splitmix64 -> (0,1) */
    if (!rng_open) { fprintf(stderr, "unif_rand() outside GetRNGstate/PutRNGstate\n"); exit(4); }
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}
void R_CheckUserInterrupt(void) {}
int R_ToplevelExec(void (*fun)(void*), void* data) { fun(data); return 1; }
static DllInfo the_dll;
int R_registerRoutines(DllInfo* d, const void* c, const R_CallMethodDef* call, const void* f, const void* e) { (void)c; (void)f; (void)e; d->call_methods = call; return 1; }
int R_useDynamicSymbols(DllInfo* d, int v) { d->dynamic_symbols = v; return 1; }

void R_init_gpirt(DllInfo*);
typedef SEXP (*call7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*call10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

int main(int argc, char** argv) {
    the_dll.dynamic_symbols = 1;
    R_init_gpirt(&the_dll);   /* what dyn.load() does for useDynLib(gpirt, .registration = TRUE) */
    if (!the_dll.call_methods || the_dll.dynamic_symbols != 0) return 2;
    const R_CallMethodDef *def = NULL, *ext = NULL;
    int count = 0;
    for (const R_CallMethodDef* p = the_dll.call_methods; p->name; ++p) {
        ++count;
        if (!strcmp(p->name, "_gpirt_gpirtMCMC")) def = p;
        if (!strcmp(p->name, "_gpirt_gpirtMCMC_b200")) ext = p;
    }
    if (!def || def->numArgs != 7 || !ext || ext->numArgs != 10 || count != 2) return 2;
    if (argc < 3) { printf("registered %s arity %d\nregistered %s arity %d\n", def->name, def->numArgs, ext->name, ext->numArgs); return 0; }
    const int thin = argc > 3 ? atoi(argv[3]) : 0;   /* > 0: call the extended routine with thin, store_f = TRUE, f_summary = TRUE */
    FILE* in = fopen(argv[1], "rb");
    if (!in) return 5;
    int hdr[4];
    if (fread(hdr, sizeof(int), 4, in) != 4 || fread(&rng_state, sizeof(rng_state), 1, in) != 1) return 5;
    const int n = hdr[0], m = hdr[1], S = hdr[2], B = hdr[3];
    SEXP y = Rf_allocMatrix(REALSXP, n, m), theta = Rf_allocVector(REALSXP, n), pm = Rf_allocMatrix(REALSXP, 2, m),
         psd = Rf_allocMatrix(REALSXP, 2, m), pstep = Rf_allocMatrix(REALSXP, 2, m);
    if (fread(REAL(y), 8, (size_t)n * m, in) != (size_t)n * m || fread(REAL(theta), 8, n, in) != (size_t)n ||
        fread(REAL(pm), 8, 2 * m, in) != (size_t)2 * m || fread(REAL(psd), 8, 2 * m, in) != (size_t)2 * m ||
        fread(REAL(pstep), 8, 2 * m, in) != (size_t)2 * m) return 5;
    fclose(in);
    SEXP sS = Rf_allocVector(REALSXP, 1), sB = Rf_allocVector(INTSXP, 1);   /* R may pass doubles or integers */
    REAL(sS)[0] = S; INTEGER(sB)[0] = B;
    uint64_t st = rng_state;   /* replicate the shim's seed derivation for the Python side to reuse */
    rng_open = 1; uint64_t lo = (uint64_t)(unif_rand() * 4294967296.0), hi = (uint64_t)(unif_rand() * 4294967296.0); rng_open = 0;
    const uint64_t seed = (hi << 32) | (lo & 0xFFFFFFFFu);
    rng_state = st;
    int jc = setjmp(error_jmp);
    if (jc) return 3;
    SEXP res;
    if (thin > 0) {
        SEXP sT = Rf_allocVector(REALSXP, 1), sF = Rf_allocVector(INTSXP, 1), sM = Rf_allocVector(INTSXP, 1);
        REAL(sT)[0] = thin; INTEGER(sF)[0] = 1; INTEGER(sM)[0] = 1;
        res = ((call10)ext->fun)(y, theta, sS, sB, pm, psd, pstep, sT, sF, sM);
    } else {
        res = ((call7)def->fun)(y, theta, sS, sB, pm, psd, pstep);
    }
    const int len = thin > 0 ? 6 : 4, slots = thin > 0 ? S / thin + 1 : S + 1;
    if (res->type != VECSXP || res->length != len || !res->names) return 6;
    const char* want[6] = {"theta", "beta", "f", "IRFs", "f_mean", "f_sd"};
    for (int i = 0; i < len; ++i) if (strcmp(CHAR(VECTOR_ELT(res->names, i)), want[i])) return 6;
    SEXP th = VECTOR_ELT(res, 0), be = VECTOR_ELT(res, 1), f = VECTOR_ELT(res, 2), irf = VECTOR_ELT(res, 3);
    if (th->dims[0] != slots || th->dims[1] != n || be->ndim != 3 || be->dims[0] != 2 || be->dims[1] != m || be->dims[2] != slots ||
        f->dims[0] != n || f->dims[1] != m || f->dims[2] != slots || irf->dims[0] != 1001 || irf->dims[1] != m) return 7;
    FILE* out = fopen(argv[2], "wb");
    fwrite(&seed, sizeof(seed), 1, out);
    fwrite(REAL(th), 8, (size_t)th->length, out); fwrite(REAL(be), 8, (size_t)be->length, out);
    fwrite(REAL(f), 8, (size_t)f->length, out); fwrite(REAL(irf), 8, (size_t)irf->length, out);
    if (thin > 0) {
        SEXP fm = VECTOR_ELT(res, 4), fsd = VECTOR_ELT(res, 5);
        if (fm->dims[0] != n || fm->dims[1] != m || fsd->dims[0] != n || fsd->dims[1] != m) return 7;
        fwrite(REAL(fm), 8, (size_t)fm->length, out); fwrite(REAL(fsd), 8, (size_t)fsd->length, out);
    }
    fclose(out);
    return 0;
}
