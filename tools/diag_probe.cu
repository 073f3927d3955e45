// Developer probe for the 128 x 128 diagonal-block kernel of the Cholesky chain (gpirt_b200/csrc/chol_diag.cuh):
// correctness against a long-double CPU factorisation, phase time stamps (clock64) and back-to-back launch time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/diag_probe tools/diag_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../gpirt_b200/csrc/chol_diag.cuh"

namespace gpirt {
void set_last_error(const char*, ...) {}
std::atomic<int64_t> g_launch_count{0};
std::mutex& device_once_mutex() { static std::mutex mu; return mu; }
}  // namespace gpirt
using namespace gpirt;

int main(int argc, char** argv) {
    const int nb = argc > 1 ? atoi(argv[1]) : 128;
    const int n = 128, ld = 136;
    std::vector<double> A((size_t)ld * n, 0.0), th(n);
    srand(7);
    for (int i = 0; i < n; ++i) {   // theta on the 0.01 grid: duplicated rows, PD only through the jitter
        double u1 = (rand() + 1.0) / (RAND_MAX + 2.0), u2 = (rand() + 1.0) / (RAND_MAX + 2.0);
        th[i] = std::round(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2) * 100.0) / 100.0;
    }
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) A[r + (size_t)c * ld] = std::exp(-0.5 * (th[r] - th[c]) * (th[r] - th[c])) + (r == c ? 1e-3 : 0.0);
    // reference factor in long double
    std::vector<long double> L((size_t)n * n, 0.0L);
    for (int c = 0; c < nb; ++c) {
        long double d = A[c + (size_t)c * ld];
        for (int k = 0; k < c; ++k) d -= L[c + (size_t)k * n] * L[c + (size_t)k * n];
        d = sqrtl(d);
        L[c + (size_t)c * n] = d;
        for (int r = c + 1; r < nb; ++r) {
            long double v = A[r + (size_t)c * ld];
            for (int k = 0; k < c; ++k) v -= L[r + (size_t)k * n] * L[c + (size_t)k * n];
            L[r + (size_t)c * n] = v / d;
        }
    }
    double *dA, *dA0, *dX; int* dst; long long* dbg;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dA0, A.size() * 8); cudaMalloc(&dX, A.size() * 8); cudaMalloc(&dst, 4); cudaMalloc(&dbg, 64 * 8);
    cudaMemcpy(dA0, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dst, 0, 4); cudaMemset(dbg, 0, 64 * 8);
    cudaFuncSetAttribute(diag::k_diag128<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag::SMEM_BYTES);
    cudaFuncSetAttribute(diag::k_diag128<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag::SMEM_BYTES);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemcpy(dA, dA0, A.size() * 8, cudaMemcpyDeviceToDevice);
        cudaMemset(dX, 0xff, A.size() * 8);
        diag::k_diag128<true><<<1, diag::DTHREADS, diag::SMEM_BYTES>>>(dA, ld, nb, dX, ld, dst, dbg);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[64];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[] = {"load", "lead0", "wait0", "syrk0", "lead1", "wait1", "syrk1", "lead2", "wait2", "syrk2", "lead3", "wait3", "-", "inv32", "offdiag32", "offdiag64"};
        printf("%s run, phases (cycles):", rep ? "warm" : "cold");
        for (int i = 0; i < 16; ++i) printf(" %s %lld |", names[i], h[i + 1] - h[i]);
        printf(" total %lld\n", h[16] - h[0]);
    }
    std::vector<double> Lg(A.size()), Xg(A.size());
    int st = 0;
    cudaMemcpy(Lg.data(), dA, A.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Xg.data(), dX, A.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
    double errL = 0, errX = 0, errU = 0;
    for (int c = 0; c < nb; ++c)
        for (int r = c; r < nb; ++r) errL = fmax(errL, fabs(Lg[r + (size_t)c * ld] - (double)L[r + (size_t)c * n]));
    for (int c = 0; c < nb; ++c)
        for (int r = 0; r < nb; ++r) {
            long double acc = 0;
            for (int k = 0; k < nb; ++k) acc += (long double)(r >= k ? Xg[r + (size_t)k * ld] : 0.0) * L[k + (size_t)c * n];
            errX = fmax(errX, fabs((double)acc - (r == c ? 1.0 : 0.0)));
            if (r < c) errU = fmax(errU, fabs(Xg[r + (size_t)c * ld]));
        }
    printf("nb %d: status %d  max|L - L_ref| %.3e  max|X L - I| %.3e  max|X upper| %.3e\n", nb, st, errL, errX, errU);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int R = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < R; ++i) diag::k_diag128<false><<<1, diag::DTHREADS, diag::SMEM_BYTES>>>(dA0, ld, nb, dX, ld, dst, nullptr);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("back-to-back launches: %.2f us per launch\n", 1000.0 * ms / R);
    return (errL < 1e-9 && errX < 1e-7 && st == 0) ? 0 : 2;
}
