// Piecewise-polynomial h(t) = log(1 + exp(-t)), t >= 0 — the per-observation term of the reference's logistic
// log-likelihood (src/log-likelihood.cpp:19-20, :33-34):   log(1 + exp(-a)) = max(-a, 0) + h(|a|).
//
// FP64 exp() + log() cost ~55 FP64-pipe operations per term on a part whose FP64 rate is 64 FMA/clk/SM; the ESS and
// beta kernels evaluate 10^8..10^9 terms per sweep.  h is analytic with small derivatives, so 2048 intervals of width
// 40/2048 with degree-5 Chebyshev interpolants reproduce it to < 2e-16 absolute (tests: vs log1p(exp(-t))) with
// 5 FMAs and three 16-byte table loads (48-byte rows, 96 KB, L1/L2 resident).  For t >= 40, h < 4.3e-18 -> 0.
#pragma once
#include "common.cuh"

namespace gpirt {

constexpr int SP_INTERVALS = 2048;
constexpr double SP_TMAX = 40.0;
constexpr int SP_DEG = 5;

// device pointer to the table (SP_INTERVALS x 6 doubles), built and uploaded on first use (softplus_table.cu)
int softplus_table(const double** dev_table);

__device__ __forceinline__ double sp_h(const double* __restrict__ tab, double t) {   // t >= 0
    if (!(t < SP_TMAX)) return (t == t) ? 0.0 : t;   // beyond the table: 0; NaN propagates
    const double s = t * (SP_INTERVALS / SP_TMAX);
    const int idx = (int)s;
    const double v = 2.0 * (s - (double)idx) - 1.0;                       // local coordinate in [-1, 1)
    const double2* row = reinterpret_cast<const double2*>(tab + (size_t)idx * 6);
    const double2 c01 = __ldg(row), c23 = __ldg(row + 1), c45 = __ldg(row + 2);
    double r = fma(c45.y, v, c45.x);
    r = fma(r, v, c23.y);
    r = fma(r, v, c23.x);
    r = fma(r, v, c01.y);
    return fma(r, v, c01.x);
}

// log(1 + exp(-a)) for any finite a; +inf where the reference's literal formula overflows (a < -709.78)
__device__ __forceinline__ double ll_term_fast(const double* __restrict__ tab, double a) {
    const double h = sp_h(tab, fabs(a));
    if (a < -709.782712893384) return INFINITY;
    return fmax(-a, 0.0) + h;
}

}  // namespace gpirt
