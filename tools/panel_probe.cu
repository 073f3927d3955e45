// Developer probe for the fused panel + next-block-column update kernel of the Cholesky chain (chol_panel.cuh): checks
// P = A21 X11^T and the rank-128 update of block column k+1 against the CPU, prints the phase time stamps of one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/panel_probe tools/panel_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../gpirt_b200/csrc/chol_panel.cuh"

namespace gpirt {
void set_last_error(const char*, ...) {}
std::atomic<int64_t> g_launch_count{0};
std::mutex& device_once_mutex() { static std::mutex mu; return mu; }
}  // namespace gpirt
using namespace gpirt;

template <int R>
static int run(int n, int k0, int probe_cta) {
    const int ld = (n + 7) / 8 * 8, r0 = k0 + 128, rem = n - r0, nb1 = rem < 128 ? rem : 128;
    std::vector<double> L((size_t)ld * n), X((size_t)ld * 128, 0.0);
    srand(3);
    for (auto& v : L) v = (rand() / (double)RAND_MAX - 0.5);
    for (int c = 0; c < 128; ++c)
        for (int r = c; r < 128; ++r) X[k0 + r + (size_t)c * ld] = (rand() / (double)RAND_MAX - 0.5) * (r == c ? 4.0 : 0.3);
    std::vector<double> ref = L;
    // P = A21 X11^T, then block column k+1 -= P P_top^T (lower triangle of its top block only)
    std::vector<double> P((size_t)rem * 128);
    for (int i = 0; i < rem; ++i)
        for (int j = 0; j < 128; ++j) {
            double a = 0;
            for (int k = 0; k <= j; ++k) a += L[r0 + i + (size_t)(k0 + k) * ld] * X[k0 + j + (size_t)k * ld];
            P[i + (size_t)j * rem] = a;
            ref[r0 + i + (size_t)(k0 + j) * ld] = a;
        }
    for (int i = 0; i < rem; ++i)
        for (int j = 0; j < nb1; ++j) {
            if (i < nb1 && j > i) continue;
            double a = 0;
            for (int k = 0; k < 128; ++k) a += P[i + (size_t)k * rem] * P[j + (size_t)k * rem];
            ref[r0 + i + (size_t)(r0 + j) * ld] -= a;
        }
    double *dL, *dX; int* cnt; long long* dbg;
    cudaMalloc(&dL, L.size() * 8); cudaMalloc(&dX, X.size() * 8); cudaMalloc(&cnt, 4); cudaMalloc(&dbg, 64 * 8);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(panel::k_panel_update<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(panel::Smem<R>));
    const int grid = (rem + R - 1) / R;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dL, L.data(), L.size() * 8, cudaMemcpyHostToDevice);
        cudaMemset(cnt, 0, 4);
        panel::k_panel_update<R, true><<<grid, panel::PTHREADS, sizeof(panel::Smem<R>)>>>(dL, ld, n, k0, dX + k0, ld, cnt, dbg, probe_cta);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[] = {"load X,A", "panel mma", "store P + fence + count", "old C + wait", "B loads issued", "update mma", "store C"};
        printf("R=%d n=%d k0=%d grid=%d cta %d rep %d:", R, n, k0, grid, probe_cta, rep);
        for (int i = 0; i < 7; ++i) printf(" %s %lld |", names[i], h[i + 1] - h[i]);
        printf(" total %lld cycles\n", h[7] - h[0]);
    }
    std::vector<double> got(L.size());
    cudaMemcpy(got.data(), dL, L.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0;
    for (size_t i = 0; i < got.size(); ++i) err = fmax(err, fabs(got[i] - ref[i]));
    printf("  max |got - ref| = %.3e\n", err);
    cudaFree(dL); cudaFree(dX); cudaFree(cnt); cudaFree(dbg);
    return err < 1e-10 ? 0 : 2;
}

int main() {
    int rc = 0;
    rc |= run<32>(4096, 0, 0);
    rc |= run<32>(4096, 0, 100);
    rc |= run<16>(2400, 128, 0);
    rc |= run<16>(2400, 128, 100);
    rc |= run<16>(517, 256, 3);
    return rc;
}
