// Host-side construction of the h(t) = log1p(exp(-t)) interpolation table (see softplus_table.cuh).
#include <cmath>
#include <mutex>
#include <vector>

#include "softplus_table.cuh"

namespace gpirt {

namespace {
// solve the (deg+1) x (deg+1) Vandermonde system in long double by Gaussian elimination with partial pivoting
void fit_interval(long double a, long double b, double* coef) {
    constexpr int P = SP_DEG + 1;
    long double V[P][P + 1];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < P; ++i) {
        const long double v = cosl(pi * (2 * i + 1) / (2.0L * P));          // Chebyshev node in [-1, 1]
        const long double t = 0.5L * (a + b) + 0.5L * (b - a) * v;
        long double p = 1.0L;
        for (int j = 0; j < P; ++j) { V[i][j] = p; p *= v; }
        V[i][P] = log1pl(expl(-t));
    }
    for (int c = 0; c < P; ++c) {
        int piv = c;
        for (int r = c + 1; r < P; ++r) if (fabsl(V[r][c]) > fabsl(V[piv][c])) piv = r;
        for (int j = 0; j <= P; ++j) std::swap(V[c][j], V[piv][j]);
        for (int r = c + 1; r < P; ++r) {
            const long double f = V[r][c] / V[c][c];
            for (int j = c; j <= P; ++j) V[r][j] -= f * V[c][j];
        }
    }
    long double x[P];
    for (int r = P - 1; r >= 0; --r) {
        long double s = V[r][P];
        for (int j = r + 1; j < P; ++j) s -= V[r][j] * x[j];
        x[r] = s / V[r][r];
    }
    for (int j = 0; j < P; ++j) coef[j] = (double)x[j];
}
}  // namespace

void softplus_table_host(std::vector<double>& tab) {
    tab.resize((size_t)SP_INTERVALS * 6);
    const long double w = (long double)SP_TMAX / SP_INTERVALS;
    for (int i = 0; i < SP_INTERVALS; ++i) fit_interval(i * w, (i + 1) * w, &tab[(size_t)i * 6]);
}

int softplus_table(const double** dev_table) {
    // one table per device, built once per process
    static std::mutex mu;
    static double* tabs[64] = {nullptr};
    int dev = 0;
    GP_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) { set_last_error("device ordinal out of range"); return GPIRT_B200_ERR_ARG; }
    if (!tabs[dev]) {
        std::vector<double> host;
        softplus_table_host(host);
        double* d = nullptr;
        GP_CUDA(cudaMalloc((void**)&d, host.size() * sizeof(double)));
        GP_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
        tabs[dev] = d;
    }
    *dev_table = tabs[dev];
    return GPIRT_B200_OK;
}

}  // namespace gpirt
