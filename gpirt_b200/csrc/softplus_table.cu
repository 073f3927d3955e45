// Host-side construction of the (h, p) node table of h(t) = log1p(exp(-t)) in long double (see softplus_table.cuh).
#include <cmath>
#include <mutex>
#include <vector>

#include "softplus_table.cuh"

namespace gpirt {

void softplus_table_host(std::vector<double>& tab) {
    tab.resize((size_t)SP_NODES * 2);
    const long double w = (long double)SP_TMAX / SP_NODES;
    for (int i = 0; i < SP_NODES; ++i) {
        const long double t = (i + 0.5L) * w;
        tab[2 * (size_t)i] = (double)log1pl(expl(-t));
        tab[2 * (size_t)i + 1] = (double)(1.0L / (1.0L + expl(t)));
    }
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
