// Blocked FP64 Cholesky + triangular solves.  Replaces arma::chol(S,"lower") -> LAPACK dpotrf
// (reference src/gpirtMCMC.cpp:17,78,97) and arma::solve(trimatl/trimatu) -> dtrtrs (src/draw-fstar.cpp:7,19).
//
// Right-looking blocked factorisation with panel width 128 (potrf_lower_rl): the 128 x 128 diagonal block is factorised
// AND inverted inside one CTA (chol_diag.cuh); the panel below it is a product with that inverse and the trailing matrix
// gets a rank-128 update, both on the DMMA GEMM, so >95% of the n^3/3 flops run on the FP64 tensor pipe.  Triangular
// solves never substitute element-wise: the base case multiplies by the stored inverse of a diagonal block (the approach
// of blocked GPU TRSMs), which is again a GEMM.
#include "chol_diag.cuh"
#include "chol_panel.cuh"
#include "gemm_f64.cuh"
#include "linalg.cuh"

#include <algorithm>

namespace gpirt {

// Inverses of the 64 x 64 diagonal blocks of an existing lower-triangular factor (stand-alone triangular solve,
// gpirt_b200_trsm_lower).  One CTA of 256 threads per block: X = L^-1 by forward substitution, one quad of lanes per
// column.  Block b = blockIdx.x works on the diagonal block starting at row/column 64 b of the n_total x n_total matrix.
__global__ void __launch_bounds__(256, 1) k_trtri64(const double* __restrict__ A, int64_t lda, int n_total,
                                                    double* __restrict__ Dinv, int64_t ldd) {
    constexpr int NB = DIAG_NB, LD = NB + 1;
    const int nb = min(NB, n_total - NB * (int)blockIdx.x);
    A += (int64_t)NB * blockIdx.x * (lda + 1);
    Dinv += (int64_t)NB * blockIdx.x;
    extern __shared__ double dsm[];
    double* Ls = dsm;                 // Ls[r * LD + c] = L(r, c)
    double* Xs = dsm + NB * LD;       // Xs[c * LD + r] = X(r, c)
    double* rinv = dsm + 2 * NB * LD; // 1 / L(r, r)
    const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;

    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx % NB, c = idx / NB;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < nb && c < nb && i >= c) v = A[i + (int64_t)c * lda];
        Ls[i * LD + c] = v;
    }
    __syncthreads();
    if (tid < NB) rinv[tid] = 1.0 / Ls[tid * LD + tid];
    __syncthreads();
    // column c handled by quad c (lanes 4c..4c+3 of one warp): x_rc = (delta_rc - sum_{k=c}^{r-1} L_rk x_kc) / L_rr
    {
        const int c = r;  // quad index = column
        for (int rr = 0; rr < nb; ++rr) {
            double part = 0.0;
            if (c < rr)
                for (int k = c + q; k < rr; k += 4) part += Ls[rr * LD + k] * Xs[c * LD + k];
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0) {
                double x = 0.0;
                if (c == rr) x = rinv[rr];
                else if (c < rr) x = -part * rinv[rr];
                Xs[c * LD + rr] = x;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx % NB, c = idx / NB;
        if (i < nb && c < nb) Dinv[i + (int64_t)c * ldd] = (i >= c) ? Xs[c * LD + i] : 0.0;
    }
}

int trtri_diag_blocks(cudaStream_t stream, const double* L, int64_t ldl, int n, double* Dinv, int64_t ldd) {
    if (n <= 0) return GPIRT_B200_OK;
    constexpr size_t smem = (size_t)(2 * DIAG_NB * (DIAG_NB + 1) + DIAG_NB) * sizeof(double);
    static bool configured[64] = {false};
    {
        DeviceOnce once(configured);
        if (once.first) GP_CUDA(cudaFuncSetAttribute(k_trtri64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    GP_LAUNCH(k_trtri64, (unsigned)ceil_div(n, DIAG_NB), 256, smem, stream, L, ldl, n, Dinv, ldd);
    GP_CUDA(cudaGetLastError());
    return GPIRT_B200_OK;
}

static int split_point(int n) {  // largest multiple of 64 that is <= half the 64-blocks (>= 64)
    const int nblk = (int)ceil_div(n, DIAG_NB);
    return (nblk / 2) * DIAG_NB;
}

static GemmArgs mk(int M, int N, int K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, double alpha, double beta, int tri) {
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.alpha = alpha; g.beta = beta; g.tri = tri;
    return g;
}

static int trsm_left_n(cudaStream_t stream, int n, int nrhs, const double* L, int64_t ldl, const double* Dinv,
                       int64_t ldd, double* B, int64_t ldb) {
    if (n <= DIAG_NB)  // B <- Linv * B  (single M tile => in place is safe)
        return gemm_f64(stream, false, false, mk(n, nrhs, n, Dinv, ldd, B, ldb, B, ldb, 1.0, 0.0, TRI_NONE));
    const int n1 = split_point(n);
    GP_TRY(trsm_left_n(stream, n1, nrhs, L, ldl, Dinv, ldd, B, ldb));
    GP_TRY(gemm_f64(stream, false, false, mk(n - n1, nrhs, n1, L + n1, ldl, B, ldb, B + n1, ldb, -1.0, 1.0, TRI_NONE)));
    return trsm_left_n(stream, n - n1, nrhs, L + n1 + (int64_t)n1 * ldl, ldl, Dinv + n1, ldd, B + n1, ldb);
}

static int trsm_left_t(cudaStream_t stream, int n, int nrhs, const double* L, int64_t ldl, const double* Dinv,
                       int64_t ldd, double* B, int64_t ldb) {
    if (n <= DIAG_NB)  // B <- Linv^T * B
        return gemm_f64(stream, true, false, mk(n, nrhs, n, Dinv, ldd, B, ldb, B, ldb, 1.0, 0.0, TRI_NONE));
    const int n1 = split_point(n);
    GP_TRY(trsm_left_t(stream, n - n1, nrhs, L + n1 + (int64_t)n1 * ldl, ldl, Dinv + n1, ldd, B + n1, ldb));
    // B1 -= L21^T X2
    GP_TRY(gemm_f64(stream, true, false, mk(n1, nrhs, n - n1, L + n1, ldl, B + n1, ldb, B, ldb, -1.0, 1.0, TRI_NONE)));
    return trsm_left_t(stream, n1, nrhs, L, ldl, Dinv, ldd, B, ldb);
}

int trsm_left_lower(cudaStream_t stream, bool trans, int n, int nrhs, const double* L, int64_t ldl,
                    const double* Dinv, int64_t ldd, double* B, int64_t ldb) {
    if (n <= 0 || nrhs <= 0) return GPIRT_B200_OK;
    return trans ? trsm_left_t(stream, n, nrhs, L, ldl, Dinv, ldd, B, ldb)
                 : trsm_left_n(stream, n, nrhs, L, ldl, Dinv, ldd, B, ldb);
}

// Right-looking Cholesky with one-panel look-ahead on two streams.
//   main stream (critical path):  diag(k) -> [wait head of bulk(k-1)] -> panel+update(k) -> diag(k+1) ...
//   aux stream  (bulk work)     :  [wait panel+update(k)] -> bulk(k) = head (block column k+2), then the rest
// panel+update(k) (chol_panel.cuh) turns the panel below diagonal block k into L and applies its rank-128 update to
// block column k+1 only; bulk(k) is the update of the remaining trailing matrix (columns >= k+2).  The critical path
// only ever waits for the one block column of a bulk update it is about to overwrite, so the serial diagonal-block
// kernels overlap the large symmetric updates and the aux stream may run up to a whole bulk update behind.
template <int R>
static int launch_panel_update(cudaStream_t stream, double* A, int64_t lda, int n, int k0, const double* Dinv, int64_t ldd,
                               int* counter) {
    constexpr size_t smem = sizeof(panel::Smem<R>);
    static bool configured[64] = {false};
    {
        DeviceOnce once(configured);
        if (once.first) GP_CUDA(cudaFuncSetAttribute(panel::k_panel_update<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int rem = n - k0 - CHOL_NB;
    GP_LAUNCH(panel::k_panel_update<R>, (unsigned)ceil_div(rem, R), panel::PTHREADS, smem, stream, A, lda, n, k0, Dinv, ldd, counter, (long long*)nullptr, 0);
    GP_CUDA(cudaGetLastError());
    return GPIRT_B200_OK;
}

int potrf_lower_rl(cudaStream_t stream, double* A, int64_t lda, int n, double* Dinv, int64_t ldd, int* d_status,
                   int* d_flags, CholLookahead* la) {
    if (n <= 0) return GPIRT_B200_OK;
    if ((ldd & 1) || (reinterpret_cast<uintptr_t>(Dinv) & 15)) {
        set_last_error("potrf_lower_rl: the block-inverse buffer must be 16-byte aligned with an even leading dimension");
        return GPIRT_B200_ERR_ARG;
    }
    constexpr size_t smem = diag::SMEM_BYTES;
    static bool configured[64] = {false};
    static int sm_count[64] = {0};
    int dev = 0;
    GP_CUDA(cudaGetDevice(&dev));
    {
        DeviceOnce once(configured);
        if (once.first) {
            GP_CUDA(cudaFuncSetAttribute(diag::k_diag128<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) GP_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
        }
    }
    const int sms = (dev >= 0 && dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
    const int nblk = (int)ceil_div(n, CHOL_NB);
    const bool two = la && la->aux && nblk > 2;
    if (two) {
        while ((int)la->ev_panel.size() < nblk) {
            cudaEvent_t e1, e2;
            GP_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            GP_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            la->ev_panel.push_back(e1); la->ev_bulk.push_back(e2);
        }
    }
    GP_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)nblk * sizeof(int), stream));   // one P_top counter per panel step
    int last_bulk = -1;
    for (int k = 0; k < nblk; ++k) {
        const int k0 = k * CHOL_NB;
        const int nb = min(CHOL_NB, n - k0);
        double* Akk = A + (int64_t)k0 * (lda + 1);
        GP_LAUNCH(diag::k_diag128<false>, 1, diag::DTHREADS, smem, stream, Akk, lda, nb, Dinv + k0, ldd, d_status, nullptr);
        GP_CUDA(cudaGetLastError());
        const int rem = n - k0 - nb;
        if (rem <= 0) {
            if (two && la->after_panel) {
                GP_CUDA(cudaEventRecord(la->ev_panel[k], stream));
                GP_TRY(la->after_panel(k, nblk, la->ev_panel[k]));
            }
            break;
        }
        // bulk(k-1) also wrote block column k+1, which panel+update(k) rewrites: only that block column of bulk(k-1) is
        // waited for (it is the first thing bulk(k-1) does), the rest of bulk(k-1) keeps running on the aux stream
        if (two && last_bulk >= 0) GP_CUDA(cudaStreamWaitEvent(stream, la->ev_bulk[last_bulk], 0));
        // 16 rows per CTA as soon as that still fits one wave of CTAs (shorter critical path), else 32
        if (ceil_div(rem, 16) <= sms) GP_TRY(launch_panel_update<16>(stream, A, lda, n, k0, Dinv + k0, ldd, d_flags + k));
        else GP_TRY(launch_panel_update<32>(stream, A, lda, n, k0, Dinv + k0, ldd, d_flags + k));
        const int nb1 = min(CHOL_NB, rem);            // width of block column k+1
        const int m2 = rem - nb1;                     // order of the trailing matrix from block column k+2 on
        const int nb2 = min(CHOL_NB, m2);             // width of block column k+2
        double* P2 = Akk + nb + nb1;                  // rows of the panel from block row k+2 on
        double* A22 = Akk + (int64_t)(nb + nb1) * (lda + 1);
        GemmArgs head, rest;                          // bulk(k) = update of block column k+2, then of columns >= k+3 (lower triangle)
        head.M = m2; head.N = nb2; head.K = nb; head.A = P2; head.lda = lda; head.B = P2; head.ldb = lda;
        head.C = A22; head.ldc = lda; head.alpha = -1.0; head.beta = 1.0; head.tri = TRI_C_LOWER;
        rest = head;
        rest.M = rest.N = m2 - nb2; rest.A = rest.B = P2 + nb2; rest.C = A22 + (int64_t)nb2 * (lda + 1);
        if (!two) {
            if (m2 > 0) GP_TRY(gemm_f64(stream, false, true, head));
            if (m2 - nb2 > 0) GP_TRY(gemm_f64(stream, false, true, rest));
            continue;
        }
        GP_CUDA(cudaEventRecord(la->ev_panel[k], stream));
        if (la->after_panel) GP_TRY(la->after_panel(k, nblk, la->ev_panel[k]));
        last_bulk = -1;
        if (m2 > 0) {
            GP_CUDA(cudaStreamWaitEvent(la->aux, la->ev_panel[k], 0));
            GP_TRY(gemm_f64(la->aux, false, true, head));
            GP_CUDA(cudaEventRecord(la->ev_bulk[k], la->aux));   // block column k+2 is up to date with panels <= k
            last_bulk = k;
            if (m2 - nb2 > 0) GP_TRY(gemm_f64(la->aux, false, true, rest));
        }
    }
    if (two) {   // join the aux stream (its last launches follow the last recorded ev_bulk)
        if ((int)la->ev_bulk.size() <= nblk) {
            cudaEvent_t e;
            GP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            la->ev_bulk.push_back(e);
        }
        GP_CUDA(cudaEventRecord(la->ev_bulk.back(), la->aux));
        GP_CUDA(cudaStreamWaitEvent(stream, la->ev_bulk.back(), 0));
    }
    return GPIRT_B200_OK;
}

__global__ void k_scatter_block_inverses(const double* __restrict__ Dinv, int64_t ldd, int n, double* __restrict__ X,
                                         int64_t ldx) {
    // X(r, c) = Dinv(r, c mod 128) inside the 128 x 128 diagonal blocks; everything else was zeroed by the caller
    const int r = blockIdx.y * blockDim.x + threadIdx.x, c = blockIdx.x;   // columns on grid.x (n may exceed 65535)
    if (r >= n || c >= n) return;
    if (r / CHOL_NB == c / CHOL_NB) X[r + (int64_t)c * ldx] = Dinv[r + (int64_t)(c % CHOL_NB) * ldd];
}

// one level-s merge of `batch` pairs starting at row/column o (pair stride 2s): X21 = -X22 (L21 X11), `rows` = order of X22
static int trtri_merge(cudaStream_t stream, const double* L, int64_t ldl, double* X, int64_t ldx, double* T, int64_t ldt, int64_t o,
                       int64_t s, int rows, int batch, double* ws, int* ws_count) {
    GemmArgs a;   // T21 = L21 X11   (X11 lower triangular)
    a.M = rows; a.N = (int)s; a.K = (int)s;
    a.A = L + (o + s) + o * ldl; a.lda = ldl; a.B = X + o + o * ldx; a.ldb = ldx;
    a.C = T + (o + s) + o * ldt; a.ldc = ldt; a.tri = TRI_B_LOWER;
    a.batch = batch;
    a.strideA = 2 * s * (ldl + 1); a.strideB = 2 * s * (ldx + 1); a.strideC = 2 * s * (ldt + 1);
    if (ws && ws_count && a.K >= 512) { a.splitk = (int)std::min<int64_t>(8, a.K / 128); a.ws = ws; a.ws_count = ws_count; }
    GP_TRY(gemm_f64(stream, false, false, a));
    GemmArgs b;   // X21 = -X22 T21  (X22 lower triangular)
    b.M = rows; b.N = (int)s; b.K = rows;
    b.A = X + (o + s) + (o + s) * ldx; b.lda = ldx; b.B = T + (o + s) + o * ldt; b.ldb = ldt;
    b.C = X + (o + s) + o * ldx; b.ldc = ldx; b.alpha = -1.0; b.tri = TRI_A_LOWER;
    b.batch = batch; b.strideA = 2 * s * (ldx + 1); b.strideB = 2 * s * (ldt + 1); b.strideC = 2 * s * (ldx + 1);
    if (ws && ws_count && b.K >= 512) { b.splitk = (int)std::min<int64_t>(8, b.K / 128); b.ws = ws; b.ws_count = ws_count; }
    return gemm_f64(stream, false, false, b);
}

int trtri_lower(cudaStream_t stream, const double* L, int64_t ldl, int n, const double* Dinv, int64_t ldd, double* X,
                int64_t ldx, double* T, int64_t ldt, int max_block, double* ws, int* ws_count) {
    if (n <= 0) return GPIRT_B200_OK;
    const int64_t stop = max_block > 0 ? std::min<int64_t>(n, max_block) : n;
    // the n x n square only: X may be a diagonal block of a larger matrix
    GP_CUDA(cudaMemset2DAsync(X, (size_t)ldx * sizeof(double), 0, (size_t)n * sizeof(double), (size_t)n, stream));
    {
        dim3 grid((unsigned)n, (unsigned)ceil_div(n, 128));
        GP_LAUNCH(k_scatter_block_inverses, grid, 128, 0, stream, Dinv, ldd, n, X, ldx);
        GP_CUDA(cudaGetLastError());
    }
    for (int64_t s = CHOL_NB; s < stop; s *= 2) {
        const int full_pairs = (int)(n / (2 * s));           // pairs whose second block is complete
        const int64_t o_r = (int64_t)full_pairs * 2 * s;     // offset of a possible ragged pair
        const int s2 = (int)std::min<int64_t>(s, n - (o_r + s));   // size of its second block (<= 0: none)
        if (full_pairs > 0) GP_TRY(trtri_merge(stream, L, ldl, X, ldx, T, ldt, 0, s, (int)s, full_pairs, ws, ws_count));
        if (s2 > 0) GP_TRY(trtri_merge(stream, L, ldl, X, ldx, T, ldt, o_r, s, s2, 1, ws, ws_count));
    }
    return GPIRT_B200_OK;
}

__global__ void k_copy_block_inverse(const double* __restrict__ Dinv, int64_t ldd, int nb, double* __restrict__ X, int64_t ldx) {
    const int r = threadIdx.x, c = blockIdx.x;   // one 128 x 128 diagonal block (strict upper of Dinv is zero)
    if (r < nb && c < nb) X[r + (int64_t)c * ldx] = Dinv[r + (int64_t)c * ldd];
}

int trtri_lower_step(cudaStream_t stream, const double* L, int64_t ldl, int n, const double* Dinv, int64_t ldd, double* X,
                     int64_t ldx, double* T, int64_t ldt, int max_block, int k, double* ws, int* ws_count) {
    const int nblk = (int)ceil_div(n, CHOL_NB);
    if (k < 0 || k >= nblk) return GPIRT_B200_OK;
    const int64_t k0 = (int64_t)k * CHOL_NB;
    const int nb = (int)std::min<int64_t>(CHOL_NB, n - k0);
    GP_LAUNCH(k_copy_block_inverse, (unsigned)nb, CHOL_NB, 0, stream, Dinv + k0, ldd, nb, X + k0 * (ldx + 1), ldx);
    GP_CUDA(cudaGetLastError());
    const int64_t stop = max_block > 0 ? std::min<int64_t>(n, max_block) : n;
    const int64_t done = k0 + nb;                        // rows / columns of L that are final
    for (int64_t s = CHOL_NB; s < stop; s *= 2) {
        if (done % (2 * s) == 0) {                       // panel k completes a full pair of order-s blocks
            GP_TRY(trtri_merge(stream, L, ldl, X, ldx, T, ldt, done - 2 * s, s, (int)s, 1, ws, ws_count));
            continue;
        }
        if (k != nblk - 1) break;                        // higher levels are not complete either
        // last panel: the ragged pair of this level, if any (its second block is shorter than s; lower levels are done)
        const int64_t o_r = (n / (2 * s)) * 2 * s;
        const int64_t s2 = std::min<int64_t>(s, n - (o_r + s));
        if (s2 > 0) GP_TRY(trtri_merge(stream, L, ldl, X, ldx, T, ldt, o_r, s, (int)s2, 1, ws, ws_count));
    }
    return GPIRT_B200_OK;
}

}  // namespace gpirt
