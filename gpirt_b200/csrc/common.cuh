// Shared host/device helpers for the gpirt_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <climits>
#include <cstdint>
#include <mutex>
#include <utility>
#include <cstdio>
#include <cstdlib>
#include <cmath>

#include "../../include/gpirt_b200.h"

namespace gpirt {

constexpr int N_GRID = GPIRT_B200_N_GRID;

void set_last_error(const char* fmt, ...);

#define GP_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ::gpirt::set_last_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                                    cudaGetErrorString(e__));                                      \
            return GPIRT_B200_ERR_CUDA;                                                            \
        }                                                                                          \
    } while (0)

#define GP_TRY(call)                                                                               \
    do {                                                                                           \
        int rc__ = (call);                                                                         \
        if (rc__ != GPIRT_B200_OK) return rc__;                                                    \
    } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Device memory comes from the stream-ordered pool (cudaMallocAsync) with the release threshold lifted, so a second
// gpirtMCMC() call in the same process re-uses the first call's memory instead of paying cudaMalloc / cudaFree again.
int pool_alloc(void** p, size_t bytes, cudaStream_t st);
void pool_free(void* p, cudaStream_t st);

// Function attributes (dynamic shared memory limits) are per device and must be set before the first launch from ANY
// host thread: `flags` is a function-local static array; the guard serialises the check and the initialisation that
// follows it in the caller's scope (`first` is true exactly once per device), so a second thread cannot launch between
// the check and the attribute call.      { DeviceOnce once(flags); if (once.first) GP_CUDA(cudaFuncSetAttribute(...)); }
std::mutex& device_once_mutex();
struct DeviceOnce {
    std::unique_lock<std::mutex> lock;
    bool* slot = nullptr;
    bool first = true;
    explicit DeviceOnce(bool (&flags)[64]) : lock(device_once_mutex()) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
        slot = &flags[dev];
        first = !*slot;
    }
    ~DeviceOnce() { if (slot) *slot = true; }
};

// launch counter (gpu_launches in bench.py): every kernel launch in this library goes through GP_LAUNCH.
// The launch carries the priority of its stream as an explicit launch attribute: a plain launch inherits it from the
// stream anyway, but a kernel node CAPTURED into a CUDA graph keeps only what the launch itself says — without it the
// replayed sweep runs the Cholesky chain at the same priority as the bulk products it is supposed to overtake.
extern std::atomic<int64_t> g_launch_count;
// priority for the launches of the calling thread instead of their stream's (INT_MIN: none): lets ONE kernel of a stream
// yield to a side stream of equal priority (the ESS beside the backward substitution, sampler.cu)
extern thread_local int g_launch_priority_override;
struct LaunchPriority {
    int saved;
    explicit LaunchPriority(int p) : saved(g_launch_priority_override) { g_launch_priority_override = p; }
    ~LaunchPriority() { g_launch_priority_override = saved; }
};
template <typename... KArgs, typename... Args>
inline void launch_with_stream_priority(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int prio = 0;
    cfg.numAttrs = 0;
    if (g_launch_priority_override != INT_MIN || (stream != nullptr && cudaStreamGetPriority(stream, &prio) == cudaSuccess)) {
        if (g_launch_priority_override != INT_MIN) prio = g_launch_priority_override;
        attr[0].id = cudaLaunchAttributePriority;
        attr[0].val.priority = prio;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define GP_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                           \
        ::gpirt::launch_with_stream_priority(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__); \
        ++::gpirt::g_launch_count;                                                                 \
    } while (0)

// ---- warp / block reductions (fixed order => deterministic) ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum over the whole block; every thread gets the same value.  `buf` = 32 doubles of shared memory that no other
// phase is touching between the two barriers (callers alternate between two buffers to save a barrier).  The warp
// partials are combined by the same shuffle tree in every warp (fixed order: deterministic, identical in all threads).
__device__ __forceinline__ double block_sum(double v, double* buf) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) buf[w] = v;
    __syncthreads();
    return warp_sum(lane < nw ? buf[lane] : 0.0);
}

// The reference's per-observation log-likelihood term, src/log-likelihood.cpp:19-20 / :33-34:
//   result -= log(1 + exp(-a)),  a = y * g          (literal form: overflows to -inf for a < -709, like the reference)
__device__ __forceinline__ double ll_term(double a) { return log(1.0 + exp(-a)); }

} // namespace gpirt
