/* TEST INFRASTRUCTURE — C-ABI doorway into the reference's own sampler sources, compiled unmodified from
 * /root/reference/src against the stand-in headers in this directory (oracle/Makefile target _ref).
 * Used to (1) pin the oracle restatement (tests/test_oracle_vs_ref.py, run in the build container) and
 * (2) time the reference's CPU path beside the GPU sampler (bench.py cpu_baseline / --impl reference).
 * Nothing here is on the product path.
 */
#include "RcppArmadillo.h"

#include <cstdint>
#include <cstring>
#include <ctime>

/* the reference's internal interface, /root/reference/src/gpirt.h:3-28 (+ ess, draw-f.cpp:21; gpirtMCMC, gpirtMCMC.cpp:5) */
arma::mat draw_f(const arma::mat& f, const arma::mat& y, const arma::mat& cholS, const arma::mat& mu);
arma::mat draw_fstar(const arma::mat& f, const arma::vec& theta, const arma::vec& theta_star, const arma::mat& L,
                     const arma::mat& mu_star);
arma::vec draw_theta(const arma::vec& theta_star, const arma::mat& y, const arma::vec& theta_prior,
                     const arma::mat& fstar, const arma::mat& mu_star);
arma::mat draw_beta(const arma::mat& beta, const arma::mat& X, const arma::mat& y, const arma::mat& f,
                    const arma::mat& prior_means, const arma::mat& prior_sds, const arma::mat& proposal_sds);
arma::mat K(const arma::vec& x1, const arma::vec& x2);
double ll(const arma::vec& f, const arma::vec& y);
double ll_bar(const arma::vec& f, const arma::vec& y, const arma::vec& mu);
arma::vec ess(const arma::vec& f, const arma::vec& y, const arma::mat& cholS, const arma::mat& mu);
Rcpp::List gpirtMCMC(const arma::mat& y, arma::vec theta, const int sample_iterations, const int burn_iterations,
                     const arma::mat& beta_prior_means, const arma::mat& beta_prior_sds,
                     const arma::mat& beta_step_sizes);

/* ---- replay tape standing in for R's global RNG ---- */
int refshim_quiet = 1;
static const double* g_vals = nullptr;
static const uint8_t* g_kinds = nullptr;
static size_t g_len = 0, g_pos = 0;
static int g_err = 0;
static double tape_pop(uint8_t want) {
    if (g_pos >= g_len) { g_err = 1; return std::nan(""); }
    if (g_kinds[g_pos] != want) g_err = 2;
    return g_vals[g_pos++];
}
/* With no tape installed (timing runs: bench.py cpu_baseline / --impl reference) a self-contained generator stands in
 * for R's: xoshiro256** uniforms in (0,1), normals by the Marsaglia polar method. */
static uint64_t g_s[4] = {0x9E3779B97F4A7C15ull, 0xBF58476D1CE4E5B9ull, 0x94D049BB133111EBull, 0x2545F4914F6CDD1Dull};
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static double xo_unif(void) {
    const uint64_t r = rotl(g_s[1] * 5, 7) * 9, t = g_s[1] << 17;
    g_s[2] ^= g_s[0]; g_s[3] ^= g_s[1]; g_s[1] ^= g_s[2]; g_s[0] ^= g_s[3]; g_s[2] ^= t; g_s[3] = rotl(g_s[3], 45);
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static double xo_norm(void) {
    static int have = 0; static double spare;
    if (have) { have = 0; return spare; }
    double u, v, q;
    do { u = 2.0 * xo_unif() - 1.0; v = 2.0 * xo_unif() - 1.0; q = u * u + v * v; } while (q >= 1.0 || q == 0.0);
    const double f = std::sqrt(-2.0 * std::log(q) / q);
    spare = v * f; have = 1;
    return u * f;
}
double refshim_norm_rand(void) { return g_vals ? tape_pop('n') : xo_norm(); }
double refshim_unif_rand(void) { return g_vals ? tape_pop('u') : xo_unif(); }

static arma::mat M(const double* p, size_t r, size_t c) { return arma::mat(p, r, c); }
static arma::vec V(const double* p, size_t n) { return arma::vec(arma::mat(p, n, 1)); }
static void out(const arma::mat& a, double* dst) { std::memcpy(dst, a.memptr(), a.n_elem * sizeof(double)); }

extern "C" {

void gpref_set_tape(const double* vals, const uint8_t* kinds, size_t len) { g_vals = vals; g_kinds = kinds; g_len = len; g_pos = 0; g_err = 0; }
void gpref_clear_tape(void) { g_vals = nullptr; g_kinds = nullptr; g_len = g_pos = 0; g_err = 0; }
void gpref_seed(uint64_t s) { for (int i = 0; i < 4; ++i) { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; g_s[i] = z ^ (z >> 31); } }
size_t gpref_tape_pos(void) { return g_pos; }
int gpref_tape_error(void) { return g_err; }
void gpref_set_quiet(int q) { refshim_quiet = q; }

void gpref_K(const double* x1, int n1, const double* x2, int n2, double* o) { out(K(V(x1, n1), V(x2, n2)), o); }
double gpref_ll(const double* f, const double* y, int n) { return ll(V(f, n), V(y, n)); }
double gpref_ll_bar(const double* f, const double* y, const double* mu, int n) { return ll_bar(V(f, n), V(y, n), V(mu, n)); }
int gpref_chol_lower(double* S, int n) {
    try { out(arma::chol(M(S, n, n), "lower"), S); } catch (...) { return 1; }
    return 0;
}
void gpref_ess(const double* f, const double* y, const double* cholS, const double* mu, int n, double* o) {
    out(ess(V(f, n), V(y, n), M(cholS, n, n), M(mu, n, 1)), o);
}
void gpref_draw_f(const double* f, const double* y, const double* cholS, const double* mu, int n, int m, double* o) {
    out(draw_f(M(f, n, m), M(y, n, m), M(cholS, n, n), M(mu, n, m)), o);
}
int gpref_draw_fstar(const double* f, const double* theta, const double* theta_star, const double* L,
                     const double* mu_star, int n, int m, int N, double* o) {
    try { out(draw_fstar(M(f, n, m), V(theta, n), V(theta_star, N), M(L, n, n), M(mu_star, N, m)), o); } catch (...) { return 1; }
    return 0;
}
void gpref_draw_theta(const double* theta_star, const double* y, const double* theta_prior, const double* fstar,
                      const double* mu_star, int n, int m, int N, double* o) {
    out(draw_theta(V(theta_star, N), M(y, n, m), V(theta_prior, N), M(fstar, N, m), M(mu_star, N, m)), o);
}
void gpref_draw_beta(const double* beta, const double* X, const double* y, const double* f, const double* pm,
                     const double* psd, const double* pstep, int n, int m, double* o) {
    out(draw_beta(M(beta, 2, m), M(X, n, 2), M(y, n, m), M(f, n, m), M(pm, 2, m), M(psd, 2, m), M(pstep, 2, m)), o);
}

/* full sampler: outputs laid out as the R arrays are (column-major): theta (S+1) x n, beta 2 x m x (S+1),
 * f n x m x (S+1), IRFs 1001 x m.  Returns 0, 1 on a thrown error (chol failure), seconds (optional) wall time. */
int gpref_mcmc(const double* y, int n, int m, const double* theta, int S, int B, const double* pm, const double* psd,
               const double* pstep, double* theta_o, double* beta_o, double* f_o, double* irf_o, double* seconds) {
    timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    try {
        Rcpp::List r = gpirtMCMC(M(y, n, m), V(theta, n), S, B, M(pm, 2, m), M(psd, 2, m), M(pstep, 2, m));
        clock_gettime(CLOCK_MONOTONIC, &t1);
        out(r.mats["theta"], theta_o);
        out(r.mats["IRFs"], irf_o);
        const arma::cube& b = r.cubes["beta"]; const arma::cube& f = r.cubes["f"];
        for (size_t s = 0; s < b.n_slices; ++s) out(b.slice(s), beta_o + s * 2 * (size_t)m);
        for (size_t s = 0; s < f.n_slices; ++s) out(f.slice(s), f_o + s * (size_t)n * m);
    } catch (...) { return 1; }
    if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    return 0;
}

} /* extern "C" */
