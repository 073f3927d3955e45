// h(t) = log(1 + exp(-t)), t >= 0 — the per-observation term of the reference's logistic log-likelihood
// (src/log-likelihood.cpp:19-20, :33-34):   log(1 + exp(-a)) = max(-a, 0) + h(|a|).
//
// FP64 exp() + log() cost ~55 FP64-pipe operations per term; the ESS and beta kernels evaluate 10^8..10^9 terms per sweep.
// Every derivative of h is a polynomial in p = 1 / (1 + e^t), with q = 1 - p:
//     h1 = -p,  h2 = pq,  h3 = -pq(q-p),  h4 = pq(1-6pq),  h5 = -pq(q-p)(1-12pq)
// So a table of (h, p) at 4096 nodes (spacing 40/4096: one 16-byte load per term, 64 KB, L1-resident) and a 5th-order
// Taylor step |dt| <= 0.0049 with the coefficients formed in registers reproduce h to 1.3e-16 absolute (remainder
// < 5e-18; tests: vs log1p(exp(-t)) at 1e-15).  An earlier version kept six Chebyshev coefficients per interval (three
// 16-byte gathers per term): ncu showed the per-item kernels bound by exactly those L1 gathers, not by FP64.
// For t >= 40, h < 4.3e-18 -> 0.
#pragma once
#include "common.cuh"

namespace gpirt {

constexpr int SP_NODES = 4096;
constexpr double SP_TMAX = 40.0;

// device pointer to the table (SP_NODES x {h(t_i), p(t_i)}, t_i = (i + 1/2) 40/4096), built on first use (softplus_table.cu)
int softplus_table(const double** dev_table);

__device__ __forceinline__ double sp_h(const double* __restrict__ tab, double t) {   // t >= 0
    if (!(t < SP_TMAX)) return (t == t) ? 0.0 : t;   // beyond the table: 0; NaN propagates
    const double s = t * (SP_NODES / SP_TMAX);
    const int idx = (int)s;
    const double d = (s - ((double)idx + 0.5)) * (SP_TMAX / SP_NODES);    // t - t_idx
    const double2 hp = __ldg(reinterpret_cast<const double2*>(tab) + idx);
    const double p = hp.y, q = 1.0 - p, u = p * q, w = q - p;
    const double uw = u * w;
    double r = fma(uw * fma(-12.0, u, 1.0) * (-1.0 / 120.0), d, u * fma(-6.0, u, 1.0) * (1.0 / 24.0));
    r = fma(r, d, uw * (-1.0 / 6.0));
    r = fma(r, d, 0.5 * u);
    r = fma(r, d, -p);
    return fma(r, d, hp.x);
}

// log(1 + exp(-a)) for any finite a; +inf where the reference's literal formula overflows (a < -709.78)
__device__ __forceinline__ double ll_term_fast(const double* __restrict__ tab, double a) {
    const double h = sp_h(tab, fabs(a));
    if (a < -709.782712893384) return INFINITY;
    return fmax(-a, 0.0) + h;
}

}  // namespace gpirt
