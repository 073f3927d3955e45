/* TEST INFRASTRUCTURE — stand-in for <RcppArmadillo.h> so that the reference's own sampler sources
 * (the .cpp files under /root/reference/src, compiled from where they lie, never copied) build in a container that has no R,
 * Rcpp or Armadillo.  It provides exactly the subset of the Armadillo / Rcpp / R-nmath surface those files use:
 *
 *   arma::mat / vec / cube / uword, .col() .row() .diag() .slice() .t() .max() .min(), (i,j) and [i] access,
 *   zeros<> ones<> regspace<> fill::zeros, chol(.,"lower"), solve(trimatl|trimatu(.), .), exp cumsum sum sqrt trans,
 *   + - * / % operators;  Rcpp::List / Named / checkUserInterrupt;  Rprintf;  R::rnorm runif dnorm plogis;  M_2PI.
 *
 * Numerical back ends are the ones Armadillo itself would call: LAPACK dpotrf / dtrtrs and BLAS dgemv / dgemm
 * (scipy's bundled OpenBLAS).  R's global RNG is replaced by a replayed tape (refshim_rng.h) so the reference
 * consumes exactly the variates the oracle restatement consumed.  Element storage carries one trailing NaN so the
 * reference's out-of-bounds read theta_star[N] (src/draw-theta.cpp:28) is deterministic instead of undefined.
 */
#ifndef REFSHIM_RCPPARMADILLO_H
#define REFSHIM_RCPPARMADILLO_H

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef M_2PI
#define M_2PI 6.283185307179586476925286766559 /* R's Rmath.h */
#endif

extern "C" {
void scipy_dpotrf_(const char*, const int*, double*, const int*, int*);
void scipy_dtrtrs_(const char*, const char*, const char*, const int*, const int*, const double*, const int*, double*,
                   const int*, int*);
void scipy_dgemv_(const char*, const int*, const int*, const double*, const double*, const int*, const double*,
                  const int*, const double*, double*, const int*);
void scipy_dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                  const int*, const double*, const int*, const double*, double*, const int*);
}

namespace arma {

typedef unsigned long long uword;
namespace fill { struct zeros_t {}; static const zeros_t zeros = zeros_t(); }

class mat;
class vec;

class mat {
  public:
    uword n_rows, n_cols, n_elem;
    std::vector<double> mem; /* n_elem + 1 doubles; the extra one is a NaN guard */

    mat() : n_rows(0), n_cols(0), n_elem(0), mem(1, std::numeric_limits<double>::quiet_NaN()) {}
    mat(uword r, uword c) { init(r, c); }
    mat(uword r, uword c, fill::zeros_t) { init(r, c); }
    mat(const double* src, uword r, uword c) { init(r, c); std::copy(src, src + n_elem, mem.begin()); }
    void init(uword r, uword c) {
        n_rows = r; n_cols = c; n_elem = r * c;
        mem.assign(n_elem + 1, 0.0);
        mem[n_elem] = std::numeric_limits<double>::quiet_NaN();
    }
    double* memptr() { return mem.data(); }
    const double* memptr() const { return mem.data(); }
    double& operator()(uword i, uword j) { return mem[j * n_rows + i]; }
    const double& operator()(uword i, uword j) const { return mem[j * n_rows + i]; }
    double& operator[](uword i) { return mem[i]; }
    const double& operator[](uword i) const { return mem[i]; }

    struct col_view; struct row_view; struct diag_view;
    col_view col(uword j);
    const vec col(uword j) const;
    row_view row(uword i);
    struct const_row_view;
    const_row_view row(uword i) const;
    diag_view diag();
    mat t() const {
        mat r(n_cols, n_rows);
        for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) r(j, i) = (*this)(i, j);
        return r;
    }
    double max() const { return *std::max_element(mem.begin(), mem.begin() + n_elem); }
    double min() const { return *std::min_element(mem.begin(), mem.begin() + n_elem); }
    mat& operator+=(const mat& o) { for (uword i = 0; i < n_elem; ++i) mem[i] += o.mem[i]; return *this; }
    mat& operator*=(double s) { for (uword i = 0; i < n_elem; ++i) mem[i] *= s; return *this; }
};

class vec : public mat {
  public:
    vec() : mat() {}
    explicit vec(uword n) : mat(n, 1) {}
    vec(const mat& m) : mat(m) { /* Col(Mat): a 1 x n row is reinterpreted as a column, like Armadillo's vector ctor */
        if (n_cols != 1) { n_rows = n_elem; n_cols = 1; }
    }
};

struct mat::col_view {
    mat& M; uword j;
    col_view& operator=(const mat& v) { std::copy(v.mem.begin(), v.mem.begin() + M.n_rows, M.mem.begin() + j * M.n_rows); return *this; }
    operator vec() const { vec r(M.n_rows); std::copy(M.mem.begin() + j * M.n_rows, M.mem.begin() + (j + 1) * M.n_rows, r.mem.begin()); return r; }
};
struct mat::row_view {
    mat& M; uword i;
    row_view& operator=(const mat& v) { for (uword j = 0; j < M.n_cols; ++j) M(i, j) = v.mem[j]; return *this; }
    vec t() const { vec r(M.n_cols); for (uword j = 0; j < M.n_cols; ++j) r[j] = M(i, j); return r; }
};
struct mat::const_row_view {
    const mat& M; uword i;
    vec t() const { vec r(M.n_cols); for (uword j = 0; j < M.n_cols; ++j) r[j] = M(i, j); return r; }
};
struct mat::diag_view {
    mat& M;
    diag_view& operator+=(double s) { uword k = std::min(M.n_rows, M.n_cols); for (uword i = 0; i < k; ++i) M(i, i) += s; return *this; }
};
inline mat::col_view mat::col(uword j) { return col_view{*this, j}; }
inline const vec mat::col(uword j) const { vec r(n_rows); std::copy(mem.begin() + j * n_rows, mem.begin() + (j + 1) * n_rows, r.mem.begin()); return r; }
inline mat::row_view mat::row(uword i) { return row_view{*this, i}; }
inline mat::const_row_view mat::row(uword i) const { return const_row_view{*this, i}; }
inline mat::diag_view mat::diag() { return diag_view{*this}; }

class cube {
  public:
    uword n_rows, n_cols, n_slices;
    std::vector<mat> slices;
    cube() : n_rows(0), n_cols(0), n_slices(0) {}
    cube(uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s), slices(s, mat(r, c)) {}
    mat& slice(uword k) { return slices[k]; }
    const mat& slice(uword k) const { return slices[k]; }
};

/* generators */
template <typename T> inline T zeros(uword n) { T r(n); return r; }
template <typename T> inline T ones(uword n) { T r(n); std::fill(r.mem.begin(), r.mem.begin() + r.n_elem, 1.0); return r; }
template <typename T> inline T regspace(double start, double delta, double end) {
    /* Armadillo internal_regspace_var_delta: N = 1 + floor((end-start)/delta); x[i] = start + T(i*delta) */
    uword N = uword(1) + uword(std::floor(double(end - start) / double(delta)));
    T r(N);
    for (uword i = 0; i < N; ++i) { volatile double step = double(i) * delta; r[i] = start + step; }
    return r;
}

/* element-wise helpers */
inline mat exp(const mat& a) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = std::exp(a.mem[i]); return r; }
inline mat sqrt(const mat& a) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = std::sqrt(a.mem[i]); return r; }
inline mat cumsum(const mat& a) { /* vectors only (the reference cumsums a vec) */
    mat r(a); for (uword i = 1; i < a.n_elem; ++i) r.mem[i] = r.mem[i - 1] + a.mem[i]; return r;
}
inline mat sum(const mat& a, int dim) {
    if (dim == 0) { mat r(1, a.n_cols); for (uword j = 0; j < a.n_cols; ++j) { double s = 0.0; for (uword i = 0; i < a.n_rows; ++i) s += a(i, j); r(0, j) = s; } return r; }
    mat r(a.n_rows, 1); for (uword i = 0; i < a.n_rows; ++i) { double s = 0.0; for (uword j = 0; j < a.n_cols; ++j) s += a(i, j); r(i, 0) = s; } return r;
}
inline mat trans(const mat& a) { return a.t(); }

inline mat operator+(const mat& a, const mat& b) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = a.mem[i] + b.mem[i]; return r; }
inline mat operator%(const mat& a, const mat& b) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = a.mem[i] * b.mem[i]; return r; }
inline mat operator*(const mat& a, double s) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = a.mem[i] * s; return r; }
inline mat operator-(const mat& a, double s) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = a.mem[i] - s; return r; }
inline mat operator/(const mat& a, double s) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = a.mem[i] / s; return r; }
inline mat operator-(double s, const mat& a) { mat r(a); for (uword i = 0; i < a.n_elem; ++i) r.mem[i] = s - a.mem[i]; return r; }

/* glue_times: BLAS dgemv for mat*vec, dgemm otherwise */
inline mat operator*(const mat& a, const mat& b) {
    const double one = 1.0, zero = 0.0; const int inc = 1;
    int m = (int)a.n_rows, k = (int)a.n_cols, n = (int)b.n_cols;
    if (a.n_cols != b.n_rows) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
    mat r(a.n_rows, b.n_cols);
    if (n == 1) scipy_dgemv_("N", &m, &k, &one, a.memptr(), &m, b.memptr(), &inc, &zero, r.memptr(), &inc);
    else scipy_dgemm_("N", "N", &m, &n, &k, &one, a.memptr(), &m, b.memptr(), &k, &zero, r.memptr(), &m);
    return r;
}

/* chol(X, "lower") : dpotrf('L'), strict upper zeroed, throws like Armadillo when not PD */
inline mat chol(const mat& X, const char* layout) {
    mat r(X);
    int n = (int)X.n_rows, info = 0;
    const char* uplo = (layout[0] == 'l') ? "L" : "U";
    scipy_dpotrf_(uplo, &n, r.memptr(), &n, &info);
    if (info != 0) throw std::runtime_error("chol(): decomposition failed");
    if (layout[0] == 'l') { for (uword j = 1; j < X.n_cols; ++j) for (uword i = 0; i < j; ++i) r(i, j) = 0.0; }
    else { for (uword j = 0; j < X.n_cols; ++j) for (uword i = j + 1; i < X.n_rows; ++i) r(i, j) = 0.0; }
    return r;
}

struct trimat_tag { mat M; bool lower; };
inline trimat_tag trimatl(const mat& M) { return trimat_tag{M, true}; }
inline trimat_tag trimatu(const mat& M) { return trimat_tag{M, false}; }
/* solve(trimat?(A), B) : LAPACK dtrtrs, as auxlib::solve_trimat_* */
inline mat solve(const trimat_tag& A, const mat& B) {
    mat r(B);
    int n = (int)A.M.n_rows, nrhs = (int)B.n_cols, info = 0;
    scipy_dtrtrs_(A.lower ? "L" : "U", "N", "N", &n, &nrhs, A.M.memptr(), &n, r.memptr(), &n, &info);
    if (info != 0) throw std::runtime_error("solve(): solution not found");
    return r;
}

} // namespace arma

/* ---- R nmath + R API bits, fed from the replay tape ---- */
double refshim_norm_rand(void);
double refshim_unif_rand(void);
extern int refshim_quiet;

namespace R {
inline double rnorm(double mu, double sigma) { return mu + sigma * refshim_norm_rand(); }            /* nmath/rnorm.c */
inline double runif(double a, double b) { return a + (b - a) * refshim_unif_rand(); }                /* nmath/runif.c */
inline double dnorm(double x, double mu, double sigma, int give_log) {                               /* nmath/dnorm.c */
    const double LN_SQRT_2PI = 0.918938533204672741780329736406, INV_SQRT_2PI = 0.398942280401432677939946059934;
    x = std::fabs((x - mu) / sigma);
    if (give_log) return -(LN_SQRT_2PI + 0.5 * x * x + std::log(sigma));
    return INV_SQRT_2PI * std::exp(-0.5 * x * x) / sigma;
}
inline double plogis(double x, double location, double scale, int lower_tail, int log_p) {           /* nmath/plogis.c */
    x = (x - location) / scale;
    x = std::exp(lower_tail ? -x : x);
    return log_p ? -std::log1p(x) : 1 / (1 + x);
}
} // namespace R

inline void Rprintf(const char* fmt, ...) {
    if (refshim_quiet) return;
    va_list ap; va_start(ap, fmt); std::vprintf(fmt, ap); va_end(ap);
}

namespace Rcpp {
inline void checkUserInterrupt() {}
struct NamedMat { std::string name; arma::mat value; };
struct NamedCube { std::string name; arma::cube value; };
inline NamedMat Named(const char* name, const arma::mat& v) { return NamedMat{name, v}; }
inline NamedCube Named(const char* name, const arma::cube& v) { return NamedCube{name, v}; }
class List {
  public:
    std::map<std::string, arma::mat> mats;
    std::map<std::string, arma::cube> cubes;
    void add(const NamedMat& e) { mats[e.name] = e.value; }
    void add(const NamedCube& e) { cubes[e.name] = e.value; }
    template <typename... Args> static List create(const Args&... args) { List l; (l.add(args), ...); return l; }
};
} // namespace Rcpp

#endif
