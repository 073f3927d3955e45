// What bounds the issue rate of tcgen05.mma.kind::i8 with operands in shared memory (B200)?  One CTA per SM issues the
// plane-pair pattern of k_dgemm_i8 (8 A planes x 8 B planes, pairs s + t <= 7, two K = 32 steps per "stage") from a
// resident stage with no loads, for several tile widths N and collector hints; prints cycles per MMA and TOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_i8_shapes tools/umma_i8_shapes.cu
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((8 * 64) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
template <int N> __host__ __device__ constexpr uint32_t idesc() { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

template <int COLL>
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t id) {
    if constexpr (COLL == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(id) : "memory");
    else if constexpr (COLL == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(id) : "memory");
    else if constexpr (COLL == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(id) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(id) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ int g_random_fill = 0;
__device__ __forceinline__ uint32_t fill_word(int i) {
    if (!g_random_fill) return 0x01010101u * (uint32_t)(i & 3);
    uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    return x & 0x7f7f7f7fu ^ ((x >> 7) & 0x40404040u) * 3u;   // bytes spread over [-64, 63]
}
constexpr int A_PLANE = 128 * 64, B_PLANE = 64 * 64, STAGE = 8 * A_PLANE + 8 * B_PLANE;

// MODE 0: 36 pairs, N = 64, collector fill/use/lastuse per A plane     (k_dgemm_i8)
// MODE 1: 36 pairs, N = 64, no collector hints
// MODE 2: 36 MMAs, N = 64, the SAME A and B every time, no hints
// MODE 3: 18 MMAs, N = 128 (two adjacent B planes as one 128-row operand), distinct A per MMA, no hints
// MODE 4: 9 MMAs, N = 256, no hints
// MODE 5: 36 pairs ordered by B plane (A changes every MMA), N = 64, no hints
template <int P, int T, int MODE>
__device__ __forceinline__ void row(uint32_t tm, uint32_t sa, uint32_t sb) {
    constexpr int LAST = 7 - P;
    constexpr int COLL = MODE != 0 ? 0 : (LAST == 0 ? 0 : (T == 0 ? 1 : (T == LAST ? 3 : 2)));
    umma<COLL>(tm + (uint32_t)((P + T) * 64), desc_sw64(sa + P * A_PLANE), desc_sw64(sb + T * B_PLANE), idesc<64>());
    if constexpr (T < LAST) row<P, T + 1, MODE>(tm, sa, sb);
}
template <int P, int MODE>
__device__ __forceinline__ void all(uint32_t tm, uint32_t sa, uint32_t sb) {
    row<P, 0, MODE>(tm, sa, sb);
    if constexpr (P < 7) all<P + 1, MODE>(tm, sa, sb);
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cycles, unsigned* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STAGE);
    const uint32_t done = smem_u32(bars);
    uint32_t* slot = (uint32_t*)(bars + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < STAGE / 4; i += 128) ((uint32_t*)smem)[i] = fill_word(i);
    if (tid == 0) { mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    long long t0 = clock64();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t sa = smem_u32(smem), sb = sa + 8 * A_PLANE;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const uint32_t a = sa + k2 * 32, b = sb + k2 * 32;
                    if constexpr (MODE == 0 || MODE == 1) all<0, MODE>(tm, a, b);
                    else if constexpr (MODE == 2) {
#pragma unroll
                        for (int j = 0; j < 36; ++j) umma<0>(tm + (uint32_t)((j & 7) * 64), desc_sw64(a), desc_sw64(b), idesc<64>());
                    } else if constexpr (MODE == 3) {
#pragma unroll
                        for (int j = 0; j < 18; ++j) umma<0>(tm + (uint32_t)((j & 3) * 128), desc_sw64(a + (j & 7) * A_PLANE), desc_sw64(b + (j % 7) * B_PLANE), idesc<128>());
                    } else if constexpr (MODE == 4) {
#pragma unroll
                        for (int j = 0; j < 9; ++j) umma<0>(tm + (uint32_t)((j & 1) * 256), desc_sw64(a + (j & 7) * A_PLANE), desc_sw64(b + (j % 5) * B_PLANE), idesc<256>());
                    } else {
#pragma unroll
                        for (int t = 0; t < 8; ++t)
#pragma unroll
                            for (int p = 0; p + t < 8; ++p)
                                umma<0>(tm + (uint32_t)((p + t) * 64), desc_sw64(a + p * A_PLANE), desc_sw64(b + t * B_PLANE), idesc<64>());
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
        }
    }
    mbar_wait(done, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tm) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (r == 0xdeadbeefu) sink[blockIdx.x] = r;
    }
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

// The k_dgemm_i8 MMA pattern (MODE 0) while a second warp streams `tma_halves` x 48 KB from global memory (L2-resident,
// a private region per SM) into a second shared-memory buffer with TMA: boxes of `row_bytes` x (8192 / row_bytes) rows.
__global__ void __launch_bounds__(128, 1) k_with_tma(const __grid_constant__ CUtensorMap tm_src, int row_bytes, int iters, int tma_halves,
                                                     long long* cycles, long long* tma_cycles, unsigned* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* tbuf = smem + STAGE;
    uint64_t* bars = (uint64_t*)(smem + 2 * STAGE);
    const uint32_t done = smem_u32(bars), full0 = smem_u32(bars + 1);
    uint32_t* slot = (uint32_t*)(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < STAGE / 4; i += 128) ((uint32_t*)smem)[i] = fill_word(i);
    if (tid == 0) { mbar_init(done, 1); mbar_init(full0, 1); mbar_init(full0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    long long t0 = clock64();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t sa = smem_u32(smem), sb = sa + 8 * A_PLANE;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) all<0, 0>(tm, sa + k2 * 32, sb + k2 * 32);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
        }
    } else if (warp == 2) {
        if (elect_one()) {
            const int rows_per_box = 8192 / row_bytes, rows_per_half = 6 * rows_per_box;
            const int row0 = blockIdx.x * 2 * rows_per_half;
            for (int it = 0; it < tma_halves; ++it) {
                const int b = it & 1;
                if (it >= 2) mbar_wait(full0 + 8 * b, ((it >> 1) - 1) & 1);
                mbar_expect_tx(full0 + 8 * b, 6 * 8192);
                for (int j = 0; j < 6; ++j)
                    tma_load_2d(smem_u32(tbuf) + b * 49152 + j * 8192, &tm_src, 0, row0 + b * rows_per_half + j * rows_per_box, full0 + 8 * b);
            }
            for (int e = tma_halves - 2; e < tma_halves; ++e)
                if (e >= 0) mbar_wait(full0 + 8 * (e & 1), (e >> 1) & 1);
            tma_cycles[blockIdx.x] = clock64() - t0;
        }
    }
    mbar_wait(done, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tm) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (r == 0xdeadbeefu) sink[blockIdx.x] = r;
    }
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void run_with_tma(int row_bytes, int tma_halves_per_stage_x100, int sms, bool shared_source) {
    static encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        fn = (encode_fn)p;
    }
    const int iters = 4000;
    const int tma_halves = (int)((long long)iters * 2 * tma_halves_per_stage_x100 / 100);
    const size_t region = 2 * 49152;
    uint8_t* src; cudaMalloc(&src, region * sms); cudaMemset(src, 1, region * sms);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)(region * (shared_source ? 1 : sms) / row_bytes)};
    const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)row_bytes, (cuuint32_t)(8192 / row_bytes)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const int smem = 2 * STAGE + 1024 + 256;
    cudaFuncSetAttribute(k_with_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long *cyc, *tc; unsigned* sink;
    cudaMalloc(&cyc, sms * sizeof(long long)); cudaMalloc(&tc, sms * sizeof(long long)); cudaMalloc(&sink, sms * sizeof(unsigned));
    cudaMemset(tc, 0, sms * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) {
        k_with_tma<<<sms, 128, smem>>>(map, row_bytes, iters, tma_halves, cyc, tc, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("k_with_tma failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    }
    long long h = 0, ht = 0; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&ht, tc, sizeof(ht), cudaMemcpyDeviceToHost);
    printf("MMA pattern + TMA %3d-byte rows, %s source, %.2f stages of TMA per MMA stage: %6.1f cycles/MMA, TMA %6.1f B/clk/SM (%lld cycles)\n",
           row_bytes, shared_source ? "one shared" : "per-SM", tma_halves_per_stage_x100 / 100.0, (double)h / ((double)iters * 72),
           ht ? (double)tma_halves * 49152 / (double)ht : 0.0, ht);
    cudaFree(src); cudaFree(cyc); cudaFree(tc); cudaFree(sink);
}

template <int MODE>
void run(const char* name, int mmas_per_kstep, int N, int sms) {
    const int smem = STAGE + 1024 + 256, iters = 40000;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* cyc; unsigned* sink;
    cudaMalloc(&cyc, sms * sizeof(long long)); cudaMalloc(&sink, sms * sizeof(unsigned));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<sms, 128, smem>>>(iters, cyc, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (r > 0 && t < best) best = t;
    }
    long long h = 0; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double mmas = (double)iters * 2 * mmas_per_kstep;
    printf("%-58s %7.1f cycles/MMA (ideal %3d)  %7.0f TOP/s\n", name, (double)h / mmas, N / 2,
           (double)sms * mmas * 2.0 * 128 * N * 32 / best * 1e-9);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("N=64, 36 plane pairs, collector::a fill/use/lastuse", 36, 64, sms);
    { int one = 1; cudaMemcpyToSymbol(g_random_fill, &one, sizeof(int)); }
    run<0>("  same, RANDOM digit data in shared memory", 36, 64, sms);
    run<4>("  N=256, distinct A per MMA, RANDOM data", 9, 256, sms);
    { int zero = 0; cudaMemcpyToSymbol(g_random_fill, &zero, sizeof(int)); }
    run<1>("N=64, 36 plane pairs, no collector hints", 36, 64, sms);
    run<5>("N=64, 36 plane pairs ordered by B plane (A changes each MMA)", 36, 64, sms);
    run<2>("N=64, same A and B every MMA, no hints", 36, 64, sms);
    run<3>("N=128, distinct A per MMA, no hints", 18, 128, sms);
    run<4>("N=256, distinct A per MMA, no hints", 9, 256, sms);
    for (int rb : {64, 128}) {
        run_with_tma(rb, 0, sms, false);
        run_with_tma(rb, 50, sms, false);
        run_with_tma(rb, 100, sms, false);
        run_with_tma(rb, 200, sms, false);   // more than the kernel needs: what the TMA path sustains beside the MMAs
        run_with_tma(rb, 100, sms, true);
    }
    return 0;
}
