// Single-CTA global load/store throughput on one SM (B200): what one latency-critical CTA can move per cycle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/st_probe tools/st_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) k(double* g, long long* out, int mode) {
    extern __shared__ __align__(128) double sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += 256) sm[i] = i;
    __syncthreads();
    long long t0, t1;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    if (mode == 0) {          // STG.64 coalesced, 64 KB
        for (int i = tid; i < 8192; i += 256) g[i] = sm[i];
    } else if (mode == 1) {   // STG.128 coalesced, 64 KB
        for (int i = tid; i < 4096; i += 256) reinterpret_cast<double2*>(g)[i] = reinterpret_cast<double2*>(sm)[i];
    } else if (mode == 2) {   // bulk async store (TMA engine) 64 KB in 8 pieces
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < 8) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 8192;" ::"l"(g + tid * 1024), "r"((unsigned)__cvta_generic_to_shared(sm + tid * 1024)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (mode == 3) {   // LDG.128 all in flight: 64 KB
        double2 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = reinterpret_cast<const double2*>(g)[tid + u * 256];
#pragma unroll
        for (int u = 0; u < 16; ++u) reinterpret_cast<double2*>(sm)[tid + u * 256] = v[u];
    } else if (mode == 4) {   // bulk async load 64 KB in 8 pieces
        __shared__ unsigned long long bar;
        const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 65536;" ::"r"(b) : "memory");
        }
        __syncthreads();
        if (tid < 8)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 8192, [%2];"
                         ::"r"((unsigned)__cvta_generic_to_shared(sm + tid * 1024)), "l"(g + tid * 1024), "r"(b) : "memory");
        unsigned ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
        } while (!ok);
    } else if (mode == 5) {   // STG.64 followed by a CTA-wide wait for the stores (membar.gl)
        for (int i = tid; i < 8192; i += 256) g[i] = sm[i];
        __threadfence();
    }
    __syncthreads();
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (tid == 0) out[mode] = t1 - t0;
}
int main() {
    double* g; long long* d;
    cudaMalloc(&g, 1 << 20); cudaMalloc(&d, 64);
    cudaMemset(g, 0, 1 << 20);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const char* names[] = {"STG.64 64KB", "STG.128 64KB", "bulk store 8x8KB", "LDG.128 64KB (16 in flight)", "bulk load 8x8KB", "STG.64 64KB + membar.gl"};
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 6; ++mode) {
            k<<<1, 256, 65536>>>(g, d, mode);
            long long h[8];
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            if (rep) printf("%-32s %6lld cycles  %.1f B/clk\n", names[mode], h[mode], 65536.0 / h[mode]);
        }
    return 0;
}
