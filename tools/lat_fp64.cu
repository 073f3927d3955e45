// dependent-issue latency of FP64 ops on one warp (B200): DFMA chain, MUFU.RCP64H, LDS round trip, BAR
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double seed, int warps_active) {
    __shared__ double sm[64];
    double x = seed + threadIdx.x * 1e-9, y = 1.000001;
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    double r = x;
#pragma unroll
    for (int i = 0; i < 64; ++i) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(r)); }
    long long t2 = clock64();
    double z = r;
#pragma unroll
    for (int i = 0; i < 64; ++i) { sm[threadIdx.x & 63] = z; z = sm[(threadIdx.x + 1) & 63] + 1.0; }
    long long t3 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) __syncthreads();
    long long t4 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; }
    if (x + r + z == 12345.678) out[4] = 1;
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(d, 1.0, threads / 32);
        long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("threads %4d: DFMA dep %.1f cyc | RCP64H dep %.1f | STS+LDS+DADD %.1f | BAR %.1f\n", threads, h[0] / 256.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0);
    }
    return 0;
}
