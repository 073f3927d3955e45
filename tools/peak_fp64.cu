// FP64 peak microbenchmarks for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate,
// m16n8k{4,8,16}.f64 lowering rate, and a cuBLAS dgemm cross-check (cross-check only; never on the product path).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/peak_fp64 tools/peak_fp64.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double seed) {
    double c[NACC][2];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double seed) {
    double c[NACC][4];
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    b[0] = seed * .5; b[1] = seed * .25;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double seed) {
    double c[NACC];
    double a = 1.0 + seed * 1e-9, b = seed * 1e-3;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exp+log throughput (the ESS likelihood inner op): log(1+exp(-a))
__global__ void k_softplus(double* out, int iters, double seed) {
    double x = seed + threadIdx.x * 1e-3, s = 0;
    for (int it = 0; it < iters; ++it) { s += log(1.0 + exp(-x)); x += 1e-4; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double4* __restrict__ a, double4* __restrict__ b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
        int threads = (wps >= 8 ? 256 : wps * 32), blocks = sms * (wps * 32 / threads);
        float ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0); });
        double flop = (double)blocks * (threads / 32) * iters * 8 * 512.0;
        printf(" \"dmma884_tflops_w%d\": %.2f,\n", wps, flop / ms * 1e-9);
    }
    {
        int threads = 256, blocks = sms * 2;
        float ms = time_ms([&] { k_dmma<16><<<blocks, threads>>>(out, iters, 1.0); });
        double flop = (double)blocks * (threads / 32) * iters * 16 * 512.0;
        printf(" \"dmma884_acc16_tflops_w16\": %.2f,\n", flop / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters, 1.0); });
        flop = (double)blocks * (threads / 32) * iters * 8 * 2048.0;
        printf(" \"dmma1688_tflops_w16\": %.2f,\n", flop / ms * 1e-9);
    }
    for (int wps = 8; wps <= 32; wps *= 2) {
        int threads = 256, blocks = sms * (wps * 32 / threads);
        float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0); });
        double flop = (double)blocks * threads * (double)iters * 8 * 2.0;
        printf(" \"dfma_tflops_w%d\": %.2f,\n", wps, flop / ms * 1e-9);
    }
    {
        int threads = 256, blocks = sms * 8;
        float ms = time_ms([&] { k_softplus<<<blocks, threads>>>(out, 2000, 0.3); });
        double ev = (double)blocks * threads * 2000.0;
        printf(" \"softplus_gevals_per_s\": %.2f,\n", ev / ms * 1e-6);
    }
    {
        size_t n = (size_t)1 << 27;  // 128Mi double4 = 4 GiB each
        double4 *a, *b; CK(cudaMalloc(&a, n * 32)); CK(cudaMalloc(&b, n * 32)); CK(cudaMemset(a, 1, n * 32));
        float ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, n); });
        printf(" \"copy_gbs\": %.1f,\n", 2.0 * n * 32 / ms * 1e-6);
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    {
        cublasHandle_t h; cublasCreate(&h);
        for (int N : {4096, 8192}) {
            double *A, *B, *C; size_t bytes = (size_t)N * N * 8;
            CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
            CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
            double al = 1, be = 0;
            float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &al, A, N, B, N, &be, C, N); }, 3);
            printf(" \"cublas_dgemm_%d_tflops\": %.2f,\n", N, 2.0 * N * N * N / ms * 1e-9);
            // sustained: back-to-back for ~2 s
            if (N == 8192) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                int reps = (int)(2000.0f / ms) + 1;
                cudaEventRecord(e0);
                for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &al, A, N, B, N, &be, C, N);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float t; cudaEventElapsedTime(&t, e0, e1);
                printf(" \"cublas_dgemm_8192_sustained_tflops\": %.2f,\n", 2.0 * N * N * N * reps / t * 1e-9);
            }
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
    printf(" \"done\": true}\n");
    return 0;
}
