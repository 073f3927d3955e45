// Dependent-issue latencies on one SM (B200) that bound the Cholesky diagonal-block kernel (chol_diag.cuh):
// DFMA chain, MUFU.RCP64H, LDS round trip, BAR, DMMA.8x8x4 accumulator chain, SHFL, STS -> __syncwarp -> LDS hand-over,
// and a release/acquire flag ping-pong between two warps of a CTA through shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lat_fp64 tools/lat_fp64.cu
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

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// one warp: DMMA accumulator chains of different widths; SHFL chain; STS -> syncwarp -> LDS chain
__global__ void k2(long long* out, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    double a = seed + lane * 1e-9, b = 0.5 * seed;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) dmma(c[0][0], c[0][1], a, b);
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { dmma(c[0][0], c[0][1], a, b); dmma(c[1][0], c[1][1], a, b); }
    long long t2 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma(c[q][0], c[q][1], a, b);
    }
    long long t3 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) dmma(c[q][0], c[q][1], a, b);
    }
    long long t4 = clock64();
    double s = a;
#pragma unroll
    for (int i = 0; i < 64; ++i) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31) + 1.0;
    long long t5 = clock64();
    double z = s;
#pragma unroll
    for (int i = 0; i < 64; ++i) { sm[lane] = z; __syncwarp(); z = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
    long long t6 = clock64();
    // DMMA feeding DFMA feeding DMMA (accumulator read-after-write by the FMA pipe)
    double w0 = 1.0, w1 = 2.0;
#pragma unroll
    for (int i = 0; i < 64; ++i) { dmma(w0, w1, a, b); w0 = fma(w0, 1.0000001, 1e-9); }
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5; out[6] = t7 - t6;
    }
    double acc = s + z + w0 + w1;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1];
    if (acc == 12345.678) out[8] = 1;
}
// two warps of one CTA pass a release/acquire sequence flag back and forth through shared memory
__global__ void k3(long long* out) {
    __shared__ int flag;
    __shared__ double payload[2];
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned fa = (unsigned)__cvta_generic_to_shared(&flag);
    long long t0 = clock64();
    double v = 0.0;
    for (int i = 0; i < 64; ++i) {
        const int want = 2 * i + w;          // warp 0 waits for even values, warp 1 for odd
        int have;
        do { asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(have) : "r"(fa) : "memory"); } while (have < want);
        v += payload[w ^ 1];
        if (lane == 0) payload[w] = v + 1.0;
        __syncwarp();
        if (lane == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(fa), "r"(want + 1) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (v == 12345.678) out[1] = 1;
}
int main() {
    long long* d; cudaMalloc(&d, 128);
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(d, 1.0, threads / 32);
        long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("threads %4d: DFMA dep %.1f cyc | RCP64H dep %.1f | STS+LDS+DADD %.1f | BAR %.1f\n", threads, h[0] / 256.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0);
    }
    {
        k2<<<1, 32>>>(d, 1.0);
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("one warp: DMMA chain x1 %.1f cyc/DMMA | x2 %.1f | x4 %.1f | x8 %.1f | SHFL.f64+DADD dep %.1f | STS+syncwarp+LDS+DADD+syncwarp %.1f | DMMA->DFMA->DMMA %.1f\n",
               h[0] / 64.0, h[1] / 128.0, h[2] / 256.0, h[3] / 512.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0);
        k2<<<1, 128>>>(d, 1.0);
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("four warps (one per SMSP): DMMA chain x1 %.1f cyc/DMMA/warp | x2 %.1f | x4 %.1f | x8 %.1f\n", h[0] / 64.0, h[1] / 128.0, h[2] / 256.0, h[3] / 512.0);
        k2<<<1, 256>>>(d, 1.0);
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("eight warps (two per SMSP): DMMA chain x1 %.1f cyc/DMMA/warp | x2 %.1f | x4 %.1f | x8 %.1f\n", h[0] / 64.0, h[1] / 128.0, h[2] / 256.0, h[3] / 512.0);
    }
    {
        k3<<<1, 64>>>(d);
        long long h[1]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        printf("release/acquire flag hop between two warps through shared memory: %.1f cyc per hop\n", h[0] / 128.0);
    }
    return 0;
}
