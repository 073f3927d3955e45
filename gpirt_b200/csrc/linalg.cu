// Blocked FP64 Cholesky + triangular solves.  Replaces arma::chol(S,"lower") -> LAPACK dpotrf
// (reference src/gpirtMCMC.cpp:17,78,97) and arma::solve(trimatl/trimatu) -> dtrtrs (src/draw-fstar.cpp:7,19).
//
// Right-looking blocked factorisation with panel width 128 (potrf_lower_rl): the 128 x 128 diagonal block is factorised
// AND inverted inside one CTA (chol_diag.cuh); the panel below it is a product with that inverse and the trailing matrix
// gets a rank-128 update, both on the DMMA GEMM, so >95% of the n^3/3 flops run on the FP64 tensor pipe.  Triangular
// solves never substitute element-wise: the base case multiplies by the stored inverse of a diagonal block (the approach
// of blocked GPU TRSMs), which is again a GEMM.
#include "chol_diag.cuh"
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
//   main stream (critical path):  diag(k) -> panel(k) -> [wait bulk(k-1)] -> crit(k) -> diag(k+1) ...
//   aux stream  (bulk work)     :  [wait panel(k)] -> bulk(k)
// crit(k) is the rank-128 update of block column k+1 only, bulk(k) that of the remaining trailing matrix (columns
// >= k+2); the serial diagonal-block kernels therefore overlap the large symmetric updates.
int potrf_lower_rl(cudaStream_t stream, double* A, int64_t lda, int n, double* Dinv, int64_t ldd, int* d_status,
                   CholLookahead* la) {
    if (n <= 0) return GPIRT_B200_OK;
    constexpr size_t smem = diag::SMEM_BYTES;
    static bool configured[64] = {false};
    {
        DeviceOnce once(configured);
        if (once.first) GP_CUDA(cudaFuncSetAttribute(diag::k_diag128<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
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
        double* P = Akk + nb;  // rem x nb panel below the diagonal block
        GemmArgs g;            // P <- P Dinv_k^T
        g.M = rem; g.N = nb; g.K = nb; g.A = P; g.lda = lda; g.B = Dinv + k0; g.ldb = ldd; g.C = P; g.ldc = lda;
        if (la && la->panel_scratch) {
            // the panel step sits on the critical path: read the source from a scratch copy so the product can use
            // the small tiles (4x the CTAs of the one-tile-wide in-place form, which must keep N in a single tile)
            GP_CUDA(cudaMemcpy2DAsync(la->panel_scratch, (size_t)la->ld_scratch * sizeof(double), P, (size_t)lda * sizeof(double),
                                      (size_t)rem * sizeof(double), (size_t)nb, cudaMemcpyDeviceToDevice, stream));
            g.A = la->panel_scratch; g.lda = la->ld_scratch;
        } else {
            g.force_big = 1;   // in place: one 128-wide column tile, each CTA rewrites only rows it read
        }
        GP_TRY(gemm_f64(stream, false, true, g));
        double* A22 = Akk + (int64_t)nb * (lda + 1);
        if (!two) {
            GemmArgs u;        // A22 -= P P^T (lower triangle)
            u.M = rem; u.N = rem; u.K = nb; u.A = P; u.lda = lda; u.B = P; u.ldb = lda;
            u.C = A22; u.ldc = lda; u.alpha = -1.0; u.beta = 1.0; u.tri = TRI_C_LOWER;
            GP_TRY(gemm_f64(stream, false, true, u));
            continue;
        }
        const int nb1 = min(CHOL_NB, rem);   // width of block column k+1
        GP_CUDA(cudaEventRecord(la->ev_panel[k], stream));
        if (la->after_panel) GP_TRY(la->after_panel(k, nblk, la->ev_panel[k]));
        if (rem - nb1 > 0) {                 // bulk(k): columns >= k+2, on the aux stream
            GP_CUDA(cudaStreamWaitEvent(la->aux, la->ev_panel[k], 0));
            GemmArgs u;
            u.M = rem - nb1; u.N = rem - nb1; u.K = nb; u.A = P + nb1; u.lda = lda; u.B = P + nb1; u.ldb = lda;
            u.C = A22 + (int64_t)nb1 * (lda + 1); u.ldc = lda; u.alpha = -1.0; u.beta = 1.0; u.tri = TRI_C_LOWER;
            GP_TRY(gemm_f64(la->aux, false, true, u));
            GP_CUDA(cudaEventRecord(la->ev_bulk[k], la->aux));
        }
        if (last_bulk >= 0) GP_CUDA(cudaStreamWaitEvent(stream, la->ev_bulk[last_bulk], 0));   // bulk(k-1) also wrote column k+1
        last_bulk = (rem - nb1 > 0) ? k : -1;
        GemmArgs c;            // crit(k): block column k+1
        c.M = rem; c.N = nb1; c.K = nb; c.A = P; c.lda = lda; c.B = P; c.ldb = lda;
        c.C = A22; c.ldc = lda; c.alpha = -1.0; c.beta = 1.0; c.tri = TRI_C_LOWER;
        GP_TRY(gemm_f64(stream, false, true, c));
    }
    if (two && last_bulk >= 0) GP_CUDA(cudaStreamWaitEvent(stream, la->ev_bulk[last_bulk], 0));
    return GPIRT_B200_OK;
}

__global__ void k_scatter_block_inverses(const double* __restrict__ Dinv, int64_t ldd, int n, double* __restrict__ X,
                                         int64_t ldx) {
    // X(r, c) = Dinv(r, c mod 128) inside the 128 x 128 diagonal blocks; everything else was zeroed by the caller
    const int r = blockIdx.y * blockDim.x + threadIdx.x, c = blockIdx.x;   // columns on grid.x (n may exceed 65535)
    if (r >= n || c >= n) return;
    if (r / CHOL_NB == c / CHOL_NB) X[r + (int64_t)c * ldx] = Dinv[r + (int64_t)(c % CHOL_NB) * ldd];
}

int trtri_lower(cudaStream_t stream, const double* L, int64_t ldl, int n, const double* Dinv, int64_t ldd, double* X,
                int64_t ldx, double* T, int64_t ldt) {
    if (n <= 0) return GPIRT_B200_OK;
    GP_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * n * sizeof(double), stream));
    {
        dim3 grid((unsigned)n, (unsigned)ceil_div(n, 128));
        GP_LAUNCH(k_scatter_block_inverses, grid, 128, 0, stream, Dinv, ldd, n, X, ldx);
        GP_CUDA(cudaGetLastError());
    }
    for (int64_t s = CHOL_NB; s < n; s *= 2) {
        const int full_pairs = (int)(n / (2 * s));           // pairs whose second block is complete
        const int64_t o_r = (int64_t)full_pairs * 2 * s;     // offset of a possible ragged pair
        const int s2 = (int)std::min<int64_t>(s, n - (o_r + s));   // size of its second block (<= 0: none)
        for (int pass = 0; pass < 2; ++pass) {
            const bool ragged = pass == 1;
            if (!ragged && full_pairs == 0) continue;
            if (ragged && s2 <= 0) continue;
            const int64_t o = ragged ? o_r : 0;
            const int rows = ragged ? s2 : (int)s;
            GemmArgs a;   // T21 = L21 X11   (X11 lower triangular)
            a.M = rows; a.N = (int)s; a.K = (int)s;
            a.A = L + (o + s) + o * ldl; a.lda = ldl; a.B = X + o + o * ldx; a.ldb = ldx;
            a.C = T + (o + s) + o * ldt; a.ldc = ldt; a.tri = TRI_B_LOWER;
            a.batch = ragged ? 1 : full_pairs;
            a.strideA = 2 * s * (ldl + 1); a.strideB = 2 * s * (ldx + 1); a.strideC = 2 * s * (ldt + 1);
            GP_TRY(gemm_f64(stream, false, false, a));
            GemmArgs b;   // X21 = -X22 T21  (X22 lower triangular)
            b.M = rows; b.N = (int)s; b.K = rows;
            b.A = X + (o + s) + (o + s) * ldx; b.lda = ldx; b.B = T + (o + s) + o * ldt; b.ldb = ldt;
            b.C = X + (o + s) + o * ldx; b.ldc = ldx; b.alpha = -1.0; b.tri = TRI_A_LOWER;
            b.batch = a.batch; b.strideA = 2 * s * (ldx + 1); b.strideB = 2 * s * (ldt + 1); b.strideC = 2 * s * (ldx + 1);
            GP_TRY(gemm_f64(stream, false, false, b));
        }
    }
    return GPIRT_B200_OK;
}

}  // namespace gpirt
