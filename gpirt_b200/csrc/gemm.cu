// Host launcher for the DMMA GEMM: picks the CTA tile (128x128 / 8 warps when that fills the 148 SMs, else 64x64 /
// 4 warps so small problems still spread over the chip) and the operand-layout instantiation.
#include "gemm_f64.cuh"

namespace gpirt {

template <int BM, int BN, int WM, int WN, bool TA, bool TB, bool BABS>
static int launch_cfg(cudaStream_t stream, const GemmArgs& g) {
    auto kern = gemm_f64_kernel<BM, BN, WM, WN, TA, TB, BABS>;
    static bool configured[64] = {false};  // per instantiation and device
    constexpr size_t smem = gemm_smem_bytes<BM, BN>();
    {
        DeviceOnce once(configured);
        if (once.first) GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid((unsigned)(ceil_div(g.N, BN) * ceil_div(g.M, BM)), (unsigned)(g.splitk > 1 ? g.splitk : 1), (unsigned)g.batch);
    GP_LAUNCH(kern, grid, dim3(WM * WN * 32), smem, stream, g);
    GP_CUDA(cudaGetLastError());
    return GPIRT_B200_OK;
}

template <bool TA, bool TB>
static int launch_layout(cudaStream_t stream, const GemmArgs& g) {
    const int64_t tiles128 = ceil_div(g.M, 128) * ceil_div(g.N, 128) * g.batch;
    if (g.b_abs) {  // only the observed-mask product of the theta step (op(B) = Y^T) takes |B|
        if constexpr (!TA && TB) {
            if (tiles128 >= 112 || g.force_big) return launch_cfg<128, 128, 4, 2, TA, TB, true>(stream, g);
            return launch_cfg<64, 64, 2, 2, TA, TB, true>(stream, g);
        } else {
            set_last_error("gemm_f64: b_abs is only instantiated for op(A)=N, op(B)=T");
            return GPIRT_B200_ERR_ARG;
        }
    }
    if (tiles128 >= 112 || g.force_big) return launch_cfg<128, 128, 4, 2, TA, TB, false>(stream, g);
    return launch_cfg<64, 64, 2, 2, TA, TB, false>(stream, g);
}

int gemm_f64(cudaStream_t stream, bool ta, bool tb, const GemmArgs& g) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return GPIRT_B200_OK;
    if (g.splitk > 1 && (!g.ws || !g.ws_count || g.tri == TRI_C_LOWER)) { set_last_error("gemm_f64: split-K needs a workspace and a full C"); return GPIRT_B200_ERR_ARG; }
    if (ta && tb) { set_last_error("gemm_f64: op(A)=T with op(B)=T is not instantiated"); return GPIRT_B200_ERR_ARG; }
    if (ta) return launch_layout<true, false>(stream, g);
    if (tb) return launch_layout<false, true>(stream, g);
    return launch_layout<false, false>(stream, g);
}

}  // namespace gpirt
