// Blocked FP64 Cholesky + triangular solves.  Replaces arma::chol(S,"lower") -> LAPACK dpotrf
// (reference src/gpirtMCMC.cpp:17,78,97) and arma::solve(trimatl/trimatu) -> dtrtrs (src/draw-fstar.cpp:7,19).
//
// Recursive (cache-oblivious) blocking: the matrix is halved at multiples of 64 until a 64 x 64 diagonal block is
// left; that block is factorised AND inverted inside one CTA (k_potrf_trtri); everything else — panel solves,
// symmetric rank-k updates, the solves' off-diagonal updates — is a large-K product on the DMMA GEMM, so >95% of
// the n^3/3 flops run on the FP64 tensor pipe.  Triangular solves never substitute: the 64 x 64 base case multiplies
// by the stored inverse of the diagonal block (the approach of blocked GPU TRSMs), which is again a GEMM.
#include "gemm_f64.cuh"
#include "linalg.cuh"

namespace gpirt {

// One CTA, 256 threads = 64 rows x 4 k-slices.  Left-looking (Crout) Cholesky of an nb x nb (nb <= 64) lower block
// held in shared memory, then X = L^-1 by forward substitution, one quad of lanes per column.
// Block b = blockIdx.x works on the diagonal block starting at row/column 64 b of the n_total x n_total matrix A.
// do_factor = 0: A already holds a Cholesky factor; only the block inverses are produced.
__global__ void __launch_bounds__(256, 1) k_potrf_trtri(double* __restrict__ A, int64_t lda, int n_total,
                                                        double* __restrict__ Dinv, int64_t ldd, int* status,
                                                        int do_factor) {
    constexpr int NB = DIAG_NB, LD = NB + 1;
    const int nb = min(NB, n_total - NB * (int)blockIdx.x);
    A += (int64_t)NB * blockIdx.x * (lda + 1);
    Dinv += (int64_t)NB * blockIdx.x;
    extern __shared__ double dsm[];
    double* Ls = dsm;                 // Ls[r * LD + c] = L(r, c)
    double* Xs = dsm + NB * LD;       // Xs[c * LD + r] = X(r, c)
    double* rinv = dsm + 2 * NB * LD; // 1 / L(r, r)
    __shared__ double s_piv;
    const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;

    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx % NB, c = idx / NB;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < nb && c < nb && i >= c) v = A[i + (int64_t)c * lda];
        Ls[i * LD + c] = v;
    }
    __syncthreads();

    if (!do_factor) {
        if (tid < nb) rinv[tid] = 1.0 / Ls[tid * LD + tid];
        __syncthreads();
    }
    for (int c = 0; do_factor && c < nb; ++c) {
        double part = 0.0;
        if (r >= c && r < nb)
            for (int k = q; k < c; k += 4) part += Ls[r * LD + k] * Ls[c * LD + k];
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        const double v = Ls[r * LD + c] - part;
        if (r == c && q == 0) {
            if (!(v > 0.0)) atomicExch(status, 1);  // not positive definite (also catches NaN)
            s_piv = sqrt(v);
        }
        __syncthreads();
        const double piv = s_piv;
        if (q == 0 && r < nb) {
            if (r == c) { Ls[r * LD + c] = piv; rinv[c] = 1.0 / piv; }
            else if (r > c) Ls[r * LD + c] = v / piv;
        }
        __syncthreads();
    }

    // X = L^-1: column c handled by quad c (lanes 4c..4c+3 of one warp): x_rc = (delta_rc - sum_{k=c}^{r-1} L_rk x_kc) / L_rr
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
        if (i < nb && c < nb) {
            if (i >= c && do_factor) A[i + (int64_t)c * lda] = Ls[i * LD + c];
            Dinv[i + (int64_t)c * ldd] = (i >= c) ? Xs[c * LD + i] : 0.0;
        }
    }
}

static int launch_diag(cudaStream_t stream, double* A, int64_t lda, int n_total, double* Dinv, int64_t ldd, int* d_status,
                       int do_factor) {
    constexpr size_t smem = (size_t)(2 * DIAG_NB * (DIAG_NB + 1) + DIAG_NB) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        GP_CUDA(cudaFuncSetAttribute(k_potrf_trtri, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    GP_LAUNCH(k_potrf_trtri, (unsigned)ceil_div(n_total, DIAG_NB), 256, smem, stream, A, lda, n_total, Dinv, ldd, d_status,
              do_factor);
    GP_CUDA(cudaGetLastError());
    return GPIRT_B200_OK;
}

int trtri_diag_blocks(cudaStream_t stream, const double* L, int64_t ldl, int n, double* Dinv, int64_t ldd) {
    if (n <= 0) return GPIRT_B200_OK;
    return launch_diag(stream, const_cast<double*>(L), ldl, n, Dinv, ldd, nullptr, 0);
}

static int split_point(int n) {  // largest multiple of 64 that is <= half the 64-blocks (>= 64)
    const int nblk = (int)ceil_div(n, DIAG_NB);
    return (nblk / 2) * DIAG_NB;
}

static GemmArgs mk(int M, int N, int K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, double alpha, double beta, int tri) {
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.alpha = alpha; g.beta = beta; g.tri = tri; g.b_abs = 0;
    return g;
}

int trsm_right_lower_t(cudaStream_t stream, int rows, int n, const double* L, int64_t ldl, const double* Dinv,
                       int64_t ldd, double* X, int64_t ldx) {
    if (rows <= 0 || n <= 0) return GPIRT_B200_OK;
    if (n <= DIAG_NB)  // X <- X * Linv^T   (single N tile => in place is safe: a CTA reads only the rows it rewrites)
        return gemm_f64(stream, false, true, mk(rows, n, n, X, ldx, Dinv, ldd, X, ldx, 1.0, 0.0, TRI_NONE));
    const int c1 = split_point(n);
    GP_TRY(trsm_right_lower_t(stream, rows, c1, L, ldl, Dinv, ldd, X, ldx));
    // X2 -= X1 * L21^T
    GP_TRY(gemm_f64(stream, false, true,
                    mk(rows, n - c1, c1, X, ldx, L + c1, ldl, X + (int64_t)c1 * ldx, ldx, -1.0, 1.0, TRI_NONE)));
    return trsm_right_lower_t(stream, rows, n - c1, L + c1 + (int64_t)c1 * ldl, ldl, Dinv + c1, ldd,
                              X + (int64_t)c1 * ldx, ldx);
}

int potrf_lower(cudaStream_t stream, double* A, int64_t lda, int n, double* Dinv, int64_t ldd, int* d_status) {
    if (n <= 0) return GPIRT_B200_OK;
    if (n <= DIAG_NB) return launch_diag(stream, A, lda, n, Dinv, ldd, d_status, 1);
    const int h = split_point(n);
    GP_TRY(potrf_lower(stream, A, lda, h, Dinv, ldd, d_status));
    double* A21 = A + h;
    double* A22 = A + h + (int64_t)h * lda;
    GP_TRY(trsm_right_lower_t(stream, n - h, h, A, lda, Dinv, ldd, A21, lda));
    // A22 -= A21 A21^T, lower triangle only
    GP_TRY(gemm_f64(stream, false, true, mk(n - h, n - h, h, A21, lda, A21, lda, A22, lda, -1.0, 1.0, TRI_C_LOWER)));
    return potrf_lower(stream, A22, lda, n - h, Dinv + h, ldd, d_status);
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

}  // namespace gpirt
