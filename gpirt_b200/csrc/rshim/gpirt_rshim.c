/* R-facing shim: the DLL a `gpirt` R package loads with useDynLib(gpirt, .registration = TRUE) (reference NAMESPACE:10).
 * It re-exports exactly what the reference's generated glue exports (src/RcppExports.cpp:16-40):
 *     SEXP _gpirt_gpirtMCMC(SEXP y, SEXP theta, SEXP sample_iterations, SEXP burn_iterations,
 *                           SEXP beta_prior_means, SEXP beta_prior_sds, SEXP beta_step_sizes)     arity 7
 *     void R_init_gpirt(DllInfo*)
 * so R/RcppExports.R:4-6 and R/gpirtMCMC.R run unchanged, and forwards to the C ABI gpirt_b200_mcmc() (CUDA).  One extra
 * routine, _gpirt_gpirtMCMC_b200 (arity 10), is the same call with the draw-storage options as trailing arguments.
 * Plain C against R's public API only (no Rcpp).  Build where R exists:  R CMD SHLIB gpirt_rshim.c -L. -lgpirt_b200 ;
 * compile- and run-checked in this repo against the stand-in headers in tests/fake_r (no R in the build image).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Random.h>
#include <R_ext/Utils.h>

#include <stdint.h>
#include <string.h>

#include "gpirt_b200.h"

static void check_interrupt_fn(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }

/* progress line exactly as the reference prints it (src/gpirtMCMC.cpp:64,83,105); the interrupt check runs inside
 * R_ToplevelExec so a pending Ctrl-C becomes a return value instead of a longjmp across live CUDA resources */
static int progress_cb(double pct, void* ctx) {
    (void)ctx;
    if (pct >= 100.0) { Rprintf("\r100.000 %% complete\n"); return 0; }
    Rprintf("\r%6.3f %% complete", pct);
    return R_ToplevelExec(check_interrupt_fn, NULL) ? 0 : 1;
}

static SEXP as_real(SEXP x) { return TYPEOF(x) == REALSXP ? x : Rf_coerceVector(x, REALSXP); }

/* the common body: thin <= 1, store_f = 1, f_summary = 0 is the reference's contract (list of 4) */
static SEXP run_mcmc(SEXP ySEXP, SEXP thetaSEXP, SEXP sample_iterationsSEXP, SEXP burn_iterationsSEXP,
                     SEXP beta_prior_meansSEXP, SEXP beta_prior_sdsSEXP, SEXP beta_step_sizesSEXP,
                     int thin, int store_f, int f_summary) {
    int nprot = 0;
    SEXP y = PROTECT(as_real(ySEXP)); ++nprot;             /* const arma::mat& y: borrowed, read-only */
    SEXP theta = PROTECT(as_real(thetaSEXP)); ++nprot;
    SEXP pm = PROTECT(as_real(beta_prior_meansSEXP)); ++nprot;
    SEXP psd = PROTECT(as_real(beta_prior_sdsSEXP)); ++nprot;
    SEXP pstep = PROTECT(as_real(beta_step_sizesSEXP)); ++nprot;
    if (!Rf_isMatrix(y)) { UNPROTECT(nprot); Rf_error("y must be a numeric matrix (a response_matrix)"); }
    const int n = Rf_nrows(y), m = Rf_ncols(y);
    const int S = Rf_asInteger(sample_iterationsSEXP), B = Rf_asInteger(burn_iterationsSEXP);   /* RcppExports.cpp:22-23 */
    if (S < 0 || B < 0 || S == NA_INTEGER || B == NA_INTEGER) { UNPROTECT(nprot); Rf_error("iteration counts must be non-negative integers"); }
    if (XLENGTH(theta) != n) { UNPROTECT(nprot); Rf_error("theta must have length nrow(y)"); }
    if (XLENGTH(pm) != 2 * (R_xlen_t)m || XLENGTH(psd) != 2 * (R_xlen_t)m || XLENGTH(pstep) != 2 * (R_xlen_t)m) {
        UNPROTECT(nprot);
        Rf_error("beta prior / step matrices must be 2 x ncol(y)");
    }
    /* Rcpp::RNGScope (RcppExports.cpp:19): read R's RNG state, draw the Philox key from it, write it back —
     * set.seed() therefore reproduces a run, and the R stream advances */
    gpirt_b200_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.device = -1;
    GetRNGstate();
    {
        const uint64_t lo = (uint64_t)(unif_rand() * 4294967296.0), hi = (uint64_t)(unif_rand() * 4294967296.0);
        opts.seed = (hi << 32) | (lo & 0xFFFFFFFFu);
    }
    PutRNGstate();

    if (thin < 1) thin = 1;
    const int slots = S / thin + 1;                        /* slot 0 = initial values, slot k = sampling iteration k * thin */
    opts.thin = thin;
    opts.skip_f_draws = store_f ? 0 : 1;
    SEXP theta_draws = PROTECT(Rf_allocMatrix(REALSXP, slots, n)); ++nprot;                /* gpirtMCMC.cpp:49 */
    SEXP beta_draws = PROTECT(Rf_alloc3DArray(REALSXP, 2, m, slots)); ++nprot;             /* :50 */
    SEXP f_draws = R_NilValue;
    if (store_f) { f_draws = PROTECT(Rf_alloc3DArray(REALSXP, n, m, slots)); ++nprot; }    /* :51 */
    SEXP irfs = PROTECT(Rf_allocMatrix(REALSXP, GPIRT_B200_N_GRID, m)); ++nprot;           /* :42 */
    SEXP f_mean = R_NilValue, f_sd = R_NilValue;
    if (f_summary) {
        f_mean = PROTECT(Rf_allocMatrix(REALSXP, n, m)); ++nprot;
        f_sd = PROTECT(Rf_allocMatrix(REALSXP, n, m)); ++nprot;
        opts.f_mean_out = REAL(f_mean); opts.f_sd_out = REAL(f_sd);
    }

    const int rc = gpirt_b200_mcmc(REAL(y), n, m, REAL(theta), S, B, REAL(pm), REAL(psd), REAL(pstep), &opts,
                                   REAL(theta_draws), REAL(beta_draws), store_f ? REAL(f_draws) : NULL, REAL(irfs), progress_cb, NULL);
    if (rc == GPIRT_B200_ERR_INTERRUPT) { UNPROTECT(nprot); Rf_onintr(); return R_NilValue; }
    if (rc != GPIRT_B200_OK) {   /* END_RCPP turns C++ exceptions into R errors (RcppExports.cpp:29) */
        UNPROTECT(nprot);
        Rf_error("%s", gpirt_b200_last_error()[0] ? gpirt_b200_last_error() : gpirt_b200_strerror(rc));
    }
    if (gpirt_b200_last_degenerate_theta() > 0)   /* the reference reads theta_star[1001] out of bounds there (draw-theta.cpp:28-33) */
        Rf_warning("%ld theta draws had a degenerate grid CDF; grid point -5 was used", (long)gpirt_b200_last_degenerate_theta());
    const int len = f_summary ? 6 : 4;
    SEXP result = PROTECT(Rf_allocVector(VECSXP, len)); ++nprot;                            /* gpirtMCMC.cpp:112-115 */
    SEXP names = PROTECT(Rf_allocVector(STRSXP, len)); ++nprot;
    SET_VECTOR_ELT(result, 0, theta_draws); SET_STRING_ELT(names, 0, Rf_mkChar("theta"));
    SET_VECTOR_ELT(result, 1, beta_draws);  SET_STRING_ELT(names, 1, Rf_mkChar("beta"));
    SET_VECTOR_ELT(result, 2, f_draws);     SET_STRING_ELT(names, 2, Rf_mkChar("f"));
    SET_VECTOR_ELT(result, 3, irfs);        SET_STRING_ELT(names, 3, Rf_mkChar("IRFs"));
    if (f_summary) {
        SET_VECTOR_ELT(result, 4, f_mean);  SET_STRING_ELT(names, 4, Rf_mkChar("f_mean"));
        SET_VECTOR_ELT(result, 5, f_sd);    SET_STRING_ELT(names, 5, Rf_mkChar("f_sd"));
    }
    Rf_setAttrib(result, R_NamesSymbol, names);
    UNPROTECT(nprot);
    return result;
}

/* the reference's entry point: seven arguments, list(theta, beta, f, IRFs)          src/RcppExports.cpp:16-30 */
SEXP _gpirt_gpirtMCMC(SEXP ySEXP, SEXP thetaSEXP, SEXP sample_iterationsSEXP, SEXP burn_iterationsSEXP,
                      SEXP beta_prior_meansSEXP, SEXP beta_prior_sdsSEXP, SEXP beta_step_sizesSEXP) {
    return run_mcmc(ySEXP, thetaSEXP, sample_iterationsSEXP, burn_iterationsSEXP, beta_prior_meansSEXP, beta_prior_sdsSEXP,
                    beta_step_sizesSEXP, 1, 1, 0);
}

/* the same call with the draw-storage options as trailing arguments (not in the reference): thin = k keeps every k-th
 * sampling iteration (arrays then hold 1 + S %/% k slots), store_f = FALSE returns f = NULL, f_summary = TRUE appends the
 * posterior mean and sd of f over ALL sampling iterations (accumulated on the GPU): list(theta, beta, f, IRFs, f_mean, f_sd) */
SEXP _gpirt_gpirtMCMC_b200(SEXP ySEXP, SEXP thetaSEXP, SEXP sample_iterationsSEXP, SEXP burn_iterationsSEXP,
                           SEXP beta_prior_meansSEXP, SEXP beta_prior_sdsSEXP, SEXP beta_step_sizesSEXP,
                           SEXP thinSEXP, SEXP store_fSEXP, SEXP f_summarySEXP) {
    const int thin = Rf_asInteger(thinSEXP), store_f = Rf_asInteger(store_fSEXP), f_summary = Rf_asInteger(f_summarySEXP);
    if (thin == NA_INTEGER || thin < 1) Rf_error("thin must be a positive integer");
    return run_mcmc(ySEXP, thetaSEXP, sample_iterationsSEXP, burn_iterationsSEXP, beta_prior_meansSEXP, beta_prior_sdsSEXP,
                    beta_step_sizesSEXP, thin, store_f != 0, f_summary != 0);
}

static const R_CallMethodDef CallEntries[] = {
    {"_gpirt_gpirtMCMC", (DL_FUNC)&_gpirt_gpirtMCMC, 7},
    {"_gpirt_gpirtMCMC_b200", (DL_FUNC)&_gpirt_gpirtMCMC_b200, 10},
    {NULL, NULL, 0}
};

void R_init_gpirt(DllInfo* dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
