/* TEST INFRASTRUCTURE — CPU oracle for the GP-IRT Gibbs sweep.  NOT part of the shipped product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * A plain C++ restatement (raw column-major arrays, OpenBLAS for the LAPACK/BLAS calls Armadillo would make)
 * of the reference sampler /root/reference/src/{gpirtMCMC,draw-f,draw-fstar,draw-theta,draw-beta,
 * covariance-function,log-likelihood}.cpp and mvnormal.h.  Every function cites the lines it follows.
 * Operation order is kept (sequential sums, log(1+exp()), 1-sqrt(), min/max CDF scaling).
 *
 * PARITY STATUS: the reference ships no golden vectors / known-answer tests for this path (its only tests
 * cover response_matrix(); tests/testthat/test_response_matrix.R).  This restatement is instead pinned against
 * the reference's own sources compiled here (oracle/_ref, see oracle/Makefile + oracle/refshim/) on identical
 * replayed random tapes — tests/test_oracle_vs_ref.py.  Without oracle/_ref it is "parity unpinned".
 *
 * The only deliberate deviation is opt-in: theta_cdf_mode = 1 subtracts max(logP) before exp() in draw_theta
 * (the reference underflows to 0/0 and reads theta_star[N] out of bounds once m is a few hundred);
 * mode 0 is the literal reference (the out-of-bounds read is replaced by a NaN sentinel).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <time.h>
#include <vector>

#include "gpo_rng.h"

extern "C" {
/* scipy-bundled OpenBLAS (Fortran ABI, 32-bit ints, hidden string lengths omitted as OpenBLAS ignores them) */
void scipy_dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
void scipy_dtrtrs_(const char* uplo, const char* trans, const char* diag, const int* n, const int* nrhs,
                   const double* a, const int* lda, double* b, const int* ldb, int* info);
void scipy_dgemv_(const char* trans, const int* m, const int* n, const double* alpha, const double* a,
                  const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy);
void scipy_dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
                  const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
                  const int* ldc);
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

using gpo::Rng;

namespace {

const double TWO_PI = 6.283185307179586476925286766559; /* R's M_2PI, draw-f.cpp:34,36 */

void gemv_n(const double* A, int rows, int cols, const double* x, double* y) {
    const double one = 1.0, zero = 0.0; const int inc = 1;
    scipy_dgemv_("N", &rows, &cols, &one, A, &rows, x, &inc, &zero, y, &inc);
}

void gemm_nn(const double* A, int m, int k, const double* B, int n, double* C) {
    const double one = 1.0, zero = 0.0;
    scipy_dgemm_("N", "N", &m, &n, &k, &one, A, &m, B, &k, &zero, C, &m);
}

} // namespace

extern "C" {

void gpo_set_blas_threads(int t) { scipy_openblas_set_num_threads(t); }
int gpo_get_blas_threads(void) { return scipy_openblas_get_num_threads(); }

/* ---- covariance-function.cpp:3-14 : K(x1,x2)[i,j] = exp(-0.5 (x1_i - x2_j)^2), n1 x n2 column-major ---- */
void gpo_K(const double* x1, int n1, const double* x2, int n2, double* out) {
    for (int j = 0; j < n2; ++j)
        for (int i = 0; i < n1; ++i) {
            double diff = x1[i] - x2[j];
            out[(size_t)j * n1 + i] = std::exp(-0.5 * diff * diff);
        }
}

/* ---- gpirtMCMC.cpp:15-17 / 76-78 / 95-97 : S = K(theta,theta); S.diag() += 0.001; chol(S,"lower") ----
 * arma::chol(.,"lower") = LAPACK dpotrf('L') with the strict upper triangle zeroed; non-PD -> error (-> R stop). */
int gpo_chol_lower(double* S, int n) {
    int info = 0;
    scipy_dpotrf_("L", &n, S, &n, &info);
    if (info != 0) return info;
    for (int j = 1; j < n; ++j)
        for (int i = 0; i < j; ++i) S[(size_t)j * n + i] = 0.0;
    return 0;
}

int gpo_build_cholS(const double* theta, int n, double* L) {
    gpo_K(theta, n, theta, n, L);
    for (int i = 0; i < n; ++i) L[(size_t)i * n + i] += 0.001;
    return gpo_chol_lower(L, n);
}

/* ---- log-likelihood.cpp:12-23 : ll(f,y) = -sum_i log(1+exp(-y_i f_i)), NaN y skipped; stride lets the caller
 * walk a matrix row without the reference's row copy (same values, same order). ---- */
double gpo_ll_strided(const double* f, size_t fstride, const double* y, size_t ystride, int n) {
    double result = 0.0;
    for (int i = 0; i < n; ++i) {
        double yi = y[i * ystride];
        if (std::isnan(yi)) continue;
        double a = yi * f[i * fstride];
        result -= std::log(1 + std::exp(-a));
    }
    return result;
}
double gpo_ll(const double* f, const double* y, int n) { return gpo_ll_strided(f, 1, y, 1, n); }

/* ---- log-likelihood.cpp:25-37 : ll_bar(f,y,mu): g = f + mu first, then as ll ---- */
double gpo_ll_bar(const double* f, const double* y, const double* mu, int n) {
    double result = 0.0;
    for (int i = 0; i < n; ++i) {
        double g = f[i] + mu[i];
        if (std::isnan(y[i])) continue;
        double a = y[i] * g;
        result -= std::log(1 + std::exp(-a));
    }
    return result;
}

/* ---- RNG handles ---- */
void* gpo_rng_keyed(uint64_t seed, int record) {
    Rng* r = new Rng(); r->kind = Rng::KEYED; r->seed = seed; r->record = record != 0; return r;
}
void* gpo_rng_tape(const double* vals, const uint8_t* kinds, size_t len) {
    Rng* r = new Rng(); r->kind = Rng::TAPE;
    r->tape_val.assign(vals, vals + len); r->tape_kind.assign(kinds, kinds + len); return r;
}
void gpo_rng_free(void* h) { delete (Rng*)h; }
void gpo_rng_set_sweep(void* h, uint32_t sweep) { ((Rng*)h)->sweep = sweep; }
/* which = 0: item streams, 1: respondent streams; len = 0 clears the map (gpo_rng.h) */
void gpo_rng_set_stream_map(void* h, int which, const uint32_t* map, size_t len) {
    std::vector<uint32_t>& m = which ? ((Rng*)h)->resp_map : ((Rng*)h)->item_map;
    m.assign(map, map + len);
}
size_t gpo_rng_tape_len(void* h) { return ((Rng*)h)->tape_val.size(); }
size_t gpo_rng_tape_pos(void* h) { return ((Rng*)h)->pos; }
int gpo_rng_error(void* h) { return ((Rng*)h)->error; }
void gpo_rng_tape_copy(void* h, double* vals, uint8_t* kinds) {
    Rng* r = (Rng*)h;
    std::memcpy(vals, r->tape_val.data(), r->tape_val.size() * sizeof(double));
    std::memcpy(kinds, r->tape_kind.data(), r->tape_kind.size());
}
double gpo_keyed_uniform(uint64_t seed, uint32_t sweep, uint32_t purpose, uint32_t stream, uint32_t idx) {
    return gpo::keyed_uniform(seed, sweep, purpose, stream, idx);
}
double gpo_keyed_normal(uint64_t seed, uint32_t sweep, uint32_t purpose, uint32_t stream, uint32_t idx) {
    return gpo::keyed_normal(seed, sweep, purpose, stream, idx);
}
void gpo_philox(uint32_t* c, uint32_t k0, uint32_t k1) { gpo::philox4x32_10(c, k0, k1); }

/* ---- mvnormal.h:4-11 : z_i = rnorm(0,1) for i ascending, returns cholS * z as a DENSE mat-vec (dgemv) ---- */
static void rmvnorm(const double* cholS, int n, Rng* rng, uint32_t purpose, uint32_t item, double* z_scratch,
                    double* out) {
    for (int i = 0; i < n; ++i) z_scratch[i] = gpo::r_rnorm(0.0, 1.0, rng->norm(purpose, item, i));
    gemv_n(cholS, n, n, z_scratch, out);
}

/* ---- draw-f.cpp:21-60 : elliptical slice sampler for one item.
 * nu (optional in/out): if nu_in != NULL it is used instead of drawing (injected-proposal parity);
 * n_prop receives the number of proposals evaluated (>= 1). ---- */
int gpo_ess(const double* f, const double* y, const double* cholS, const double* mu, int n, uint32_t item,
            void* rng_h, const double* nu_in, double* f_out, double* nu_out, int* n_prop) {
    Rng* rng = (Rng*)rng_h;
    std::vector<double> nu(n), z(n);
    if (nu_in) std::copy(nu_in, nu_in + n, nu.begin());
    else rmvnorm(cholS, n, rng, gpo::P_ESS_Z, item, z.data(), nu.data());           /* :26 */
    if (nu_out) std::copy(nu.begin(), nu.end(), nu_out);
    double u = gpo::r_runif(0.0, 1.0, rng->unif(gpo::P_ESS_U, item, 0));             /* :28 */
    double log_y = gpo_ll_bar(f, y, mu, n) + std::log(u);                            /* :29 */
    double epsilon_min = 0.0;                                                       /* :33 */
    double epsilon_max = TWO_PI;                                                    /* :34 */
    double epsilon = gpo::r_runif(epsilon_min, epsilon_max, rng->unif(gpo::P_ESS_U, item, 1)); /* :35 */
    epsilon_min = epsilon - TWO_PI;                                                 /* :36 (max stays 2pi) */
    int iter = 0;
    const int ITER_CAP = 100000; /* the reference loops forever on a NaN likelihood; the oracle gives up loudly */
    for (;;) {
        iter += 1;
        double c = std::cos(epsilon), s = std::sin(epsilon);
        for (int i = 0; i < n; ++i) f_out[i] = f[i] * c + nu[i] * s;                 /* :43 */
        if (gpo_ll_bar(f_out, y, mu, n) > log_y) break;                              /* :45 strict > */
        if (epsilon < 0.0) epsilon_min = epsilon; else epsilon_max = epsilon;        /* :50-55 */
        epsilon = gpo::r_runif(epsilon_min, epsilon_max, rng->unif(gpo::P_ESS_U, item, 1 + iter)); /* :56 */
        if (iter >= ITER_CAP) { if (n_prop) *n_prop = iter; return -1; }
    }
    if (n_prop) *n_prop = iter;
    return 0;
}

/* ---- draw-f.cpp:64-73 : serial loop of ess over items.  n_prop (optional, length m). ---- */
int gpo_draw_f(const double* f, const double* y, const double* cholS, const double* mu, int n, int m,
               void* rng_h, double* f_out, int* n_prop) {
    for (int j = 0; j < m; ++j) {
        int np = 0;
        int rc = gpo_ess(f + (size_t)j * n, y + (size_t)j * n, cholS, mu + (size_t)j * n, n, (uint32_t)j, rng_h,
                         nullptr, f_out + (size_t)j * n, nullptr, &np);
        if (n_prop) n_prop[j] = np;
        if (rc) return rc;
    }
    return 0;
}

/* ---- draw-fstar.cpp:10-31 (+ double_solve :3-8).
 * kstar = K(theta,theta*) (n x N); tmp = L^-1 kstar (dtrtrs L,N); s = 1 - sqrt(colsum(tmp%tmp));
 * per item: alpha = solve(trimatu(L^T), solve(trimatl(L), f_j)); mean = kstar^T alpha + mu*_j;
 * f*_kj = rnorm(mean_k, s_k), k ascending.  Optional dumps: s_out (N), mean_out (N x m). ---- */
int gpo_draw_fstar(const double* f, const double* theta, const double* theta_star, const double* L,
                   const double* mu_star, int n, int m, int N, void* rng_h, double* fstar_out, double* s_out,
                   double* mean_out) {
    Rng* rng = (Rng*)rng_h;
    std::vector<double> kstar((size_t)n * N), kstarT((size_t)N * n), tmp, Lt((size_t)n * n), s(N), alpha(n), mean(N);
    gpo_K(theta, n, theta_star, N, kstar.data());                                   /* :17 */
    for (int k = 0; k < N; ++k) for (int i = 0; i < n; ++i) kstarT[(size_t)i * N + k] = kstar[(size_t)k * n + i]; /* :18 */
    tmp = kstar;
    int info = 0, one = 1;
    scipy_dtrtrs_("L", "N", "N", &n, &N, L, &n, tmp.data(), &n, &info);             /* :19 */
    if (info) return info;
    for (int k = 0; k < N; ++k) {                                                   /* :20 */
        double acc = 0.0;
        for (int i = 0; i < n; ++i) { double t = tmp[(size_t)k * n + i]; acc += t * t; }
        s[k] = 1.0 - std::sqrt(acc);
    }
    if (s_out) std::copy(s.begin(), s.end(), s_out);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) Lt[(size_t)i * n + j] = L[(size_t)j * n + i];  /* L.t(), :7 */
    for (int j = 0; j < m; ++j) {
        std::copy(f + (size_t)j * n, f + (size_t)(j + 1) * n, alpha.begin());
        scipy_dtrtrs_("L", "N", "N", &n, &one, L, &n, alpha.data(), &n, &info);     /* solve(trimatl(L), X) */
        if (info) return info;
        scipy_dtrtrs_("U", "N", "N", &n, &one, Lt.data(), &n, alpha.data(), &n, &info); /* solve(trimatu(L.t()), .) */
        if (info) return info;
        gemv_n(kstarT.data(), N, n, alpha.data(), mean.data());                     /* :25 */
        for (int k = 0; k < N; ++k) mean[k] = mean[k] + mu_star[(size_t)j * N + k];
        if (mean_out) std::copy(mean.begin(), mean.end(), mean_out + (size_t)j * N);
        for (int k = 0; k < N; ++k)                                                 /* :26-28 */
            fstar_out[(size_t)j * N + k] = gpo::r_rnorm(mean[k], s[k], rng->norm(gpo::P_FSTAR_Z, (uint32_t)j, k));
    }
    return 0;
}

/* ---- draw-theta.cpp:3-37.  mode 0 = literal (exp of raw log-posterior; no P_k > u -> NaN sentinel replaces the
 * reference's out-of-bounds theta_star[N]); mode 1 = max-subtracted before exp (same CDF wherever mode 0 is finite).
 * idx_out (optional) receives the chosen grid index (N = sentinel); logp_out (optional) the n x N log-posteriors. ---- */
void gpo_draw_theta(const double* theta_star, const double* y, const double* theta_prior, const double* fstar,
                    int n, int m, int N, int mode, void* rng_h, double* theta_out, int* idx_out, double* logp_out) {
    Rng* rng = (Rng*)rng_h;
    std::vector<double> P(N);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < N; ++k)                                                 /* :15-19 */
            P[k] = theta_prior[k] + gpo_ll_strided(fstar + k, (size_t)N, y + i, (size_t)n, m);
        if (logp_out) for (int k = 0; k < N; ++k) logp_out[(size_t)k * n + i] = P[k];
        if (mode == 1) {
            double mx = *std::max_element(P.begin(), P.end());
            for (int k = 0; k < N; ++k) P[k] -= mx;
        }
        for (int k = 0; k < N; ++k) P[k] = std::exp(P[k]);                           /* :21 */
        for (int k = 1; k < N; ++k) P[k] = P[k - 1] + P[k];                          /* :22 cumsum */
        double max_p = *std::max_element(P.begin(), P.end());                       /* :23 */
        double min_p = *std::min_element(P.begin(), P.end());                       /* :24 */
        for (int k = 0; k < N; ++k) P[k] = (P[k] - min_p) / (max_p - min_p);         /* :25 */
        double u = gpo::r_runif(0.0, 1.0, rng->unif(gpo::P_THETA_U, (uint32_t)i, 0)); /* :27 */
        double res = std::numeric_limits<double>::quiet_NaN();                      /* :28 theta_star[N] is OOB */
        int idx = N;
        for (int k = 0; k < N; ++k)                                                 /* :29-34 */
            if (P[k] > u) { res = theta_star[k]; idx = k; break; }
        theta_out[i] = res;
        if (idx_out) idx_out[i] = idx;
    }
}

/* ---- draw-beta.cpp:3-41 : per item, sequentially k = 0,1: RW proposal, normal prior, ll_bar ratio.
 * X is n x 2 (ones, theta).  X*pv is a dgemv as Armadillo would issue.  accept_out optional (2 x m). ---- */
void gpo_draw_beta(const double* beta, const double* X, const double* y, const double* f, const double* prior_means,
                   const double* prior_sds, const double* step_sizes, int n, int m, void* rng_h, double* beta_out,
                   int* accept_out) {
    Rng* rng = (Rng*)rng_h;
    const int p = 2;
    std::vector<double> mu_pv(n), mu_cv(n);
    for (int j = 0; j < m; ++j) {
        const double* responses = y + (size_t)j * n;
        const double* rho = f + (size_t)j * n;
        double cv[2] = {beta[2 * j], beta[2 * j + 1]};
        double pv[2] = {cv[0], cv[1]};
        for (int k = 0; k < p; ++k) {
            pv[k] = gpo::r_rnorm(cv[k], step_sizes[2 * j + k], rng->norm(gpo::P_BETA_Z, (uint32_t)j, k)); /* :22 */
            double prior_mean = prior_means[2 * j + k], prior_sd = prior_sds[2 * j + k];
            double pv_prior = gpo::r_dnorm_log(pv[k], prior_mean, prior_sd);        /* :25 */
            double cv_prior = gpo::r_dnorm_log(cv[k], prior_mean, prior_sd);        /* :26 */
            gemv_n(X, n, 2, pv, mu_pv.data());
            gemv_n(X, n, 2, cv, mu_cv.data());
            double pv_ll = gpo_ll_bar(rho, responses, mu_pv.data(), n);             /* :27 */
            double cv_ll = gpo_ll_bar(rho, responses, mu_cv.data(), n);             /* :28 */
            double r = pv_prior + pv_ll - cv_prior - cv_ll;                         /* :29 */
            int acc = std::log(gpo::r_runif(0.0, 1.0, rng->unif(gpo::P_BETA_U, (uint32_t)j, k))) < r; /* :30 */
            if (acc) cv[k] = pv[k]; else pv[k] = cv[k];
            if (accept_out) accept_out[2 * j + k] = acc;
        }
        beta_out[2 * j] = cv[0]; beta_out[2 * j + 1] = cv[1];
    }
}

/* grid + prior, gpirtMCMC.cpp:35-47: theta*_k = -5 + k*0.01 (arma::regspace: start + i*delta, N = 1+floor(10/0.01)
 * = 1001), prior_k = dnorm(theta*_k, 0, 1, log) */
int gpo_grid(double* theta_star, double* theta_prior) {
    const double start = -5.0, delta = 0.01, end = 5.0;
    int N = 1 + (int)std::floor((end - start) / delta);
    for (int i = 0; i < N; ++i) {
        volatile double step = (double)i * delta; /* volatile: forbid FMA contraction, as Armadillo's T(i*delta) */
        theta_star[i] = start + step;
        if (theta_prior) theta_prior[i] = gpo::r_dnorm_log(theta_star[i], 0.0, 1.0);
    }
    return N;
}

/* linear mean, gpirtMCMC.cpp:33,40,74-75: mu = X * beta via dgemm (n x 2 . 2 x m) */
void gpo_linear_mean(const double* x, int n, const double* beta, int m, double* mu) {
    std::vector<double> X((size_t)n * 2);
    for (int i = 0; i < n; ++i) { X[i] = 1.0; X[(size_t)n + i] = x[i]; }
    gemm_nn(X.data(), n, 2, beta, m, mu);
}

/* ---- gpirtMCMC.cpp:5-117 : the driver.  Outputs (column-major, caller-allocated):
 *   theta_draws (S+1) x n ; beta_draws 2 x m x (S+1) ; f_draws n x m x (S+1) ; irfs N x m.
 * fstar_last (optional, N x m): the final f* state, for tests.  Returns 0, or <0 on Cholesky/ESS failure.
 * rng sweep counter: 0 during init, t = 1.. for sweep t (burn-in and sampling share one counter). ---- */
int gpo_mcmc(const double* y, int n, int m, const double* theta_init, int sample_iterations, int burn_iterations,
             const double* beta_prior_means, const double* beta_prior_sds, const double* beta_step_sizes,
             void* rng_h, int theta_cdf_mode, double* theta_draws, double* beta_draws, double* f_draws,
             double* irfs, double* fstar_last, double* step_seconds /* optional [7]: f,fstar,theta,beta,mean,K+chol,total */) {
    Rng* rng = (Rng*)rng_h;
    const int N = 1001;
    const size_t nm = (size_t)n * m, Nm = (size_t)N * m;
    int total_iterations = sample_iterations + burn_iterations;
    std::vector<double> theta(theta_init, theta_init + n), cholS((size_t)n * n), f(nm), fnew(nm), beta(2 * (size_t)m),
        bnew(2 * (size_t)m), X((size_t)n * 2), mu(nm), theta_star(N), theta_prior(N), mu_star(Nm), f_star(Nm), z(n);
    rng->sweep = 0;
    if (gpo_build_cholS(theta.data(), n, cholS.data())) return -1;                  /* :15-17 */
    for (int j = 0; j < m; ++j) rmvnorm(cholS.data(), n, rng, gpo::P_INIT_F_Z, (uint32_t)j, z.data(), &f[(size_t)j * n]); /* :19-21 */
    for (int j = 0; j < m; ++j)                                                     /* :23-27 */
        for (int p = 0; p < 2; ++p)
            beta[2 * j + p] = gpo::r_rnorm(beta_prior_means[2 * j + p], beta_prior_sds[2 * j + p],
                                           rng->norm(gpo::P_INIT_BETA, (uint32_t)j, p));
    gpo_linear_mean(theta.data(), n, beta.data(), m, mu.data());                    /* :30-33 */
    gpo_grid(theta_star.data(), theta_prior.data());                                /* :35-36, :44-47 */
    gpo_linear_mean(theta_star.data(), N, beta.data(), m, mu_star.data());          /* :37-40 */
    if (gpo_draw_fstar(f.data(), theta.data(), theta_star.data(), cholS.data(), mu_star.data(), n, m, N, rng,
                       f_star.data(), nullptr, nullptr)) return -2;                 /* :41 */
    std::fill(irfs, irfs + Nm, 0.0);                                                /* :42 */
    const int S1 = sample_iterations + 1;
    for (int i = 0; i < n; ++i) theta_draws[(size_t)i * S1] = theta[i];             /* :53 */
    std::copy(beta.begin(), beta.end(), beta_draws);                                /* :54 */
    std::copy(f.begin(), f.end(), f_draws);                                         /* :55 */
    double secs[7] = {0, 0, 0, 0, 0, 0, 0};
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    double t_begin = now();
    for (int iter = 0; iter < total_iterations; ++iter) {                           /* :60-79 and :81-104 */
        rng->sweep = (uint32_t)(iter + 1);
        double t0 = now();
        if (gpo_draw_f(f.data(), y, cholS.data(), mu.data(), n, m, rng, fnew.data(), nullptr)) return -3; /* :68/:87 */
        f.swap(fnew);
        double t1 = now();
        if (gpo_draw_fstar(f.data(), theta.data(), theta_star.data(), cholS.data(), mu_star.data(), n, m, N, rng,
                           f_star.data(), nullptr, nullptr)) return -2;             /* :69/:88 */
        double t2 = now();
        gpo_draw_theta(theta_star.data(), y, theta_prior.data(), f_star.data(), n, m, N, theta_cdf_mode, rng,
                       theta.data(), nullptr, nullptr);                             /* :70/:89 */
        double t3 = now();
        for (int i = 0; i < n; ++i) { X[i] = 1.0; X[(size_t)n + i] = theta[i]; }    /* :71/:90 */
        gpo_draw_beta(beta.data(), X.data(), y, f.data(), beta_prior_means, beta_prior_sds, beta_step_sizes, n, m,
                      rng, bnew.data(), nullptr);                                   /* :72/:91 */
        beta.swap(bnew);
        double t4 = now();
        gpo_linear_mean(theta.data(), n, beta.data(), m, mu.data());                /* :74/:93 */
        gpo_linear_mean(theta_star.data(), N, beta.data(), m, mu_star.data());      /* :75/:94 */
        double t5 = now();
        if (gpo_build_cholS(theta.data(), n, cholS.data())) return -1;              /* :76-78/:95-97 */
        double t6 = now();
        secs[0] += t1 - t0; secs[1] += t2 - t1; secs[2] += t3 - t2; secs[3] += t4 - t3; secs[4] += t5 - t4; secs[5] += t6 - t5;
        if (iter >= burn_iterations) {
            int sidx = iter - burn_iterations + 1;
            for (int i = 0; i < n; ++i) theta_draws[(size_t)i * S1 + sidx] = theta[i]; /* :99 */
            std::copy(beta.begin(), beta.end(), beta_draws + (size_t)sidx * 2 * m);   /* :100 */
            std::copy(f.begin(), f.end(), f_draws + (size_t)sidx * nm);               /* :101 */
            for (size_t q = 0; q < Nm; ++q) irfs[q] += f_star[q];                     /* :103 */
        }
    }
    secs[6] = now() - t_begin;
    double scale = 1.0 / (double)sample_iterations;                                 /* :106 (S = 0 -> inf, NaN IRFs) */
    for (size_t q = 0; q < Nm; ++q) irfs[q] *= scale;
    for (size_t q = 0; q < Nm; ++q) irfs[q] = gpo::r_plogis(irfs[q]);               /* :107-111 */
    if (fstar_last) std::copy(f_star.begin(), f_star.end(), fstar_last);
    if (step_seconds) std::copy(secs, secs + 7, step_seconds);
    return 0;
}

} /* extern "C" */
