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

#include <algorithm>

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
    static bool configured[64] = {false};
    if (first_use_on_device(configured))
        GP_CUDA(cudaFuncSetAttribute(k_potrf_trtri, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    g.alpha = alpha; g.beta = beta; g.tri = tri;
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

// ======================================================================================================================
// 128 x 128 diagonal block: Cholesky factor AND its inverse in one CTA (512 threads).
//
// One shared array S[128][129] holds everything: strictly below the diagonal L, on the diagonal 1/L_rr (= the inverse's
// diagonal), strictly above the diagonal the transpose of X = L^-1 (X(r,c) lives at S[c][r]); L's own diagonal goes to
// ldiag[].  Recursion 128 -> 64 -> 32: a 32 x 32 diagonal block is factorised and inverted by ONE warp without block
// barriers (lane = row for the Crout factorisation, lane = column for the forward substitutions); the off-diagonal
// work ( L21 = A21 X11^T,  A22 -= L21 L21^T,  X21 = -X22 L21 X11 ) is done by the whole CTA from shared memory.
// ======================================================================================================================
namespace {
constexpr int DB = CHOL_NB, DLD = DB + 1, DTHREADS = 512;

// fast 1/p for p > 0: hardware double-precision reciprocal seed (MUFU.RCP64H, ~20 bits) + two Newton steps
// (4 dependent DFMAs instead of the full IEEE division routine; result within 1 ulp)
__device__ __forceinline__ double fast_rcp(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    r = fma(r, fma(-p, r, 1.0), r);
    r = fma(r, fma(-p, r, 1.0), r);
    return r;
}

// 32 x 32 diagonal block at offset o, all 512 threads, ONE barrier per column, factor and inverse in the same loop.
// Symmetric Gaussian elimination  A = Lh D Lh^T  (Lh unit lower):  step c uses the unscaled column c and the pivot p_c,
//   a_ij -= (a_ic / p_c) a_jc                         (i >= j > c)
//   W_ij -= (a_ic / p_c) W_cj ,  W_ic = -(a_ic / p_c)  (i > c >= j;  W accumulates Lh^-1, W_cc = 1 implicit)
// and afterwards  L = Lh D^1/2 : l_ij = a_ij / sqrt(p_j),   X = L^-1 = D^-1/2 W : x_ij = W_ij / sqrt(p_i).
// W_ij lives transposed at S[j][i] (strict upper triangle), pivots in pv[].
__device__ void diag32_factor_invert(double* S, double* ldiag, double* pv, int o, int tid, int* status,
                                     long long* dbg = nullptr) {
    const int j = tid & 31, i0 = tid >> 5;   // column j, rows i0 and i0 + 16
    double* pinv_s = pv + DB;                // reciprocals of the pivots, published by the thread that produces the pivot
    if (tid == 0) { const double p0 = S[o * DLD + o]; pv[o] = p0; pinv_s[o] = fast_rcp(p0); }
    __syncthreads();
    // Critical path of one step: load 1/p_c -> multiplier -> update -> (owner of a_{c+1,c+1}: reciprocal of the new
    // pivot) -> barrier.  Everything that does not depend on 1/p_c is loaded before it.
    for (int c = 0; c < 32; ++c) {
        const int cc = o + c;
        const bool probe = dbg && o == 0 && c == 8 && tid == 64;
        if (probe) dbg[10] = clock64();
        // one predicated path for all lanes (no divergence): j > c updates A(i,j); j <= c updates W(i,j) stored at
        // S[j][i], where j == c is the fresh column W(i,c) = 0 - mlt * 1 (the strict upper triangle starts as zeros)
        const double other = (j == c) ? 1.0 : S[(o + j) * DLD + cc];   // a_jc, or W(c,j) which lives at S[o+j][o+c]
        double* tgt[2];
        double mv[2], tv[2];
        bool active[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + 16 * u;
            tgt[u] = (j > c) ? &S[(o + i) * DLD + o + j] : &S[(o + j) * DLD + o + i];
            active[u] = (i > c) && (j <= c || i >= j);
            mv[u] = S[(o + i) * DLD + cc];
            tv[u] = *tgt[u];
        }
        const double pinv = pinv_s[cc];
        if (probe) dbg[11] = clock64() + (long long)(pinv == 123.456);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!active[u]) continue;
            const double r = fma(-(mv[u] * pinv), other, tv[u]);
            *tgt[u] = r;
            if (j == c + 1 && i0 + 16 * u == c + 1) { pv[cc + 1] = r; pinv_s[cc + 1] = fast_rcp(r); }   // next pivot
        }
        if (probe) dbg[12] = clock64();
        __syncthreads();
        if (probe) dbg[13] = clock64();
    }
    if (tid < 32 && !(pv[o + tid] > 0.0)) atomicExch(status, 1);   // not positive definite (or NaN)
    // scale: l_ij = a_ij rsqrt(p_j) (i > j);  x_ij = W_ij rsqrt(p_i), stored at S[j][i];  diagonals
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = i0 + 16 * u;
        if (i > j) {
            S[(o + i) * DLD + o + j] *= rsqrt(pv[o + j]);
            S[(o + j) * DLD + o + i] *= rsqrt(pv[o + i]);
        } else if (i == j) {
            const double pp = pv[o + i], rs = rsqrt(pp);
            ldiag[o + i] = pp * rs;
            S[(o + i) * DLD + o + i] = rs;
        }
    }
    __syncthreads();
}

// ---- off-diagonal block products of the recursion, on the FP64 tensor pipe straight out of shared memory ----------
// H x H blocks as 8 x 8 DMMA fragments (m8n8k4): lane (g = lane/4, t = lane%4) supplies A(row0+g, k0+t) and
// B(k0+t, col0+g) and owns C(row0+g, col0+2t..2t+1).  16 warps: H = 64 -> warp w does row block w%8 and the four
// column blocks (w/8)*4.., H = 32 -> one fragment per warp.  Operands that live transposed / triangular in the shared
// array are fetched through their index formulas, structural zeros are skipped k-step-wise (uniform per warp).
template <int H> struct FragMap {
    static constexpr int NF = H / 8, FPW = (NF * NF) / 16;
    __device__ static int rb(int w) { return w % NF; }
    __device__ static int cb(int w, int q) { return (w / NF) * FPW + q; }
};

// L21 = A21 X11^T  (block at rows ro.., cols co..; X11 = inverse of the diagonal block at co, X11(j,k) at S[co+k][co+j], k <= j)
template <int H>
__device__ void mm_panel(double* S, int ro, int co, int tid) {
    using FM = FragMap<H>;
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = FM::rb(w) * 8;
    double acc[FM::FPW][2];
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) acc[q][0] = acc[q][1] = 0.0;
    const int cmax = FM::cb(w, FM::FPW - 1) * 8 + 7;          // largest column of this warp: k runs to it
    for (int k0 = 0; k0 <= cmax; k0 += 4) {
        const double av = S[(ro + r0 + g) * DLD + co + k0 + t];
#pragma unroll
        for (int q = 0; q < FM::FPW; ++q) {
            const int c0 = FM::cb(w, q) * 8;
            if (k0 > c0 + 7) continue;                          // X11(j,k) = 0 for k > j
            const int j = c0 + g, k = k0 + t;
            const double bv = (k <= j) ? S[(co + k) * DLD + co + j] : 0.0;
            dmma_8x8x4(acc[q][0], acc[q][1], av, bv);
        }
    }
    __syncthreads();                                            // everyone has read A21 before it is overwritten
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) {
        const int c0 = FM::cb(w, q) * 8;
        S[(ro + r0 + g) * DLD + co + c0 + 2 * t] = acc[q][0];
        S[(ro + r0 + g) * DLD + co + c0 + 2 * t + 1] = acc[q][1];
    }
    __syncthreads();
}

// A22 -= L21 L21^T (lower triangle incl. diagonal); A22 at (ro, ro), L21 at (ro, co)
template <int H>
__device__ void mm_syrk(double* S, int ro, int co, int tid) {
    using FM = FragMap<H>;
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int rbk = FM::rb(w), r0 = rbk * 8;
    double acc[FM::FPW][2];
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) acc[q][0] = acc[q][1] = 0.0;
    if (FM::cb(w, 0) <= rbk) {                                  // at least one fragment on or below the diagonal
        for (int k0 = 0; k0 < H; k0 += 4) {
            const double av = S[(ro + r0 + g) * DLD + co + k0 + t];
#pragma unroll
            for (int q = 0; q < FM::FPW; ++q) {
                const int cbk = FM::cb(w, q);
                if (cbk > rbk) continue;
                const double bv = S[(ro + cbk * 8 + g) * DLD + co + k0 + t];   // L21(j, k)
                dmma_8x8x4(acc[q][0], acc[q][1], av, bv);
            }
        }
#pragma unroll
        for (int q = 0; q < FM::FPW; ++q) {
            const int cbk = FM::cb(w, q);
            if (cbk > rbk) continue;
            const int i = r0 + g, j = cbk * 8 + 2 * t;
            if (i >= j) S[(ro + i) * DLD + ro + j] -= acc[q][0];
            if (i >= j + 1) S[(ro + i) * DLD + ro + j + 1] -= acc[q][1];
        }
    }
    __syncthreads();
}

// X21 = -X22 (L21 X11); X21(i,j) is stored at S[co+j][ro+i]
template <int H>
__device__ void mm_inv_offdiag(double* S, int ro, int co, int tid) {
    using FM = FragMap<H>;
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = FM::rb(w) * 8;
    double acc[FM::FPW][2];
    // T = L21 X11 into the X21 slot:  T(i,j) = sum_{k >= j} L21(i,k) X11(k,j),  X11(k,j) = S[co+j][co+k]
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) acc[q][0] = acc[q][1] = 0.0;
    const int cmin = FM::cb(w, 0) * 8;                          // smallest column of this warp: k starts at its fragment
    for (int k0 = (cmin / 4) * 4; k0 < H; k0 += 4) {
        const double av = S[(ro + r0 + g) * DLD + co + k0 + t];
#pragma unroll
        for (int q = 0; q < FM::FPW; ++q) {
            const int c0 = FM::cb(w, q) * 8;
            if (k0 + 3 < c0) continue;                          // X11(k,j) = 0 for k < j
            const int j = c0 + g, k = k0 + t;
            const double bv = (k >= j) ? S[(co + j) * DLD + co + k] : 0.0;
            dmma_8x8x4(acc[q][0], acc[q][1], av, bv);
        }
    }
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) {
        const int c0 = FM::cb(w, q) * 8;
        S[(co + c0 + 2 * t) * DLD + ro + r0 + g] = acc[q][0];
        S[(co + c0 + 2 * t + 1) * DLD + ro + r0 + g] = acc[q][1];
    }
    __syncthreads();
    // X21(i,j) = - sum_{k <= i} X22(i,k) T(k,j),  X22(i,k) = S[ro+k][ro+i],  T(k,j) = S[co+j][ro+k]
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) acc[q][0] = acc[q][1] = 0.0;
    for (int k0 = 0; k0 <= r0 + 7; k0 += 4) {                   // X22(i,k) = 0 for k > i
        const int i = r0 + g, k = k0 + t;
        const double av = (k <= i) ? S[(ro + k) * DLD + ro + i] : 0.0;
#pragma unroll
        for (int q = 0; q < FM::FPW; ++q) {
            const int c0 = FM::cb(w, q) * 8;
            const double bv = S[(co + c0 + g) * DLD + ro + k];
            dmma_8x8x4(acc[q][0], acc[q][1], av, bv);
        }
    }
    __syncthreads();                                            // T fully consumed before X21 replaces it
#pragma unroll
    for (int q = 0; q < FM::FPW; ++q) {
        const int c0 = FM::cb(w, q) * 8;
        S[(co + c0 + 2 * t) * DLD + ro + r0 + g] = -acc[q][0];
        S[(co + c0 + 2 * t + 1) * DLD + ro + r0 + g] = -acc[q][1];
    }
    __syncthreads();
}

__device__ void factor_invert_64(double* S, double* ldiag, double* pv, int o, int tid, int* status) {
    diag32_factor_invert(S, ldiag, pv, o, tid, status);
    mm_panel<32>(S, o + 32, o, tid);
    mm_syrk<32>(S, o + 32, o, tid);
    diag32_factor_invert(S, ldiag, pv, o + 32, tid, status);
    mm_inv_offdiag<32>(S, o + 32, o, tid);
}
}  // namespace

#define DIAG_MARK(slot) do { if (dbg && tid == 0) dbg[slot] = clock64(); } while (0)
__global__ void __launch_bounds__(DTHREADS, 1) k_diag128(double* __restrict__ A, int64_t lda, int nb,
                                                         double* __restrict__ Dinv, int64_t ldd, int* status,
                                                         long long* dbg) {
    extern __shared__ double dsm[];
    double* S = dsm;
    double* ldiag = dsm + DB * DLD;
    double* pv = ldiag + DB;
    const int tid = threadIdx.x;
    DIAG_MARK(0);
    // load the lower triangle (rows/cols >= nb padded with the identity)
    for (int idx = tid; idx < DB * DB; idx += DTHREADS) {
        const int r = idx % DB, c = idx / DB;
        double v = 0.0;
        if (r >= c) v = (r < nb) ? A[r + (int64_t)c * lda] : (r == c ? 1.0 : 0.0);
        S[r * DLD + c] = v;
    }
    __syncthreads();
    DIAG_MARK(1);
    diag32_factor_invert(S, ldiag, pv, 0, tid, status, dbg);
    DIAG_MARK(2);
    mm_panel<32>(S, 32, 0, tid);
    mm_syrk<32>(S, 32, 0, tid);
    DIAG_MARK(3);
    diag32_factor_invert(S, ldiag, pv, 32, tid, status);
    mm_inv_offdiag<32>(S, 32, 0, tid);
    DIAG_MARK(4);
    mm_panel<64>(S, 64, 0, tid);
    DIAG_MARK(5);
    mm_syrk<64>(S, 64, 0, tid);
    DIAG_MARK(6);
    factor_invert_64(S, ldiag, pv, 64, tid, status);
    DIAG_MARK(7);
    mm_inv_offdiag<64>(S, 64, 0, tid);
    DIAG_MARK(8);
    for (int idx = tid; idx < DB * DB; idx += DTHREADS) {
        const int r = idx % DB, c = idx / DB;
        if (r < nb && c < nb) {
            if (r > c) A[r + (int64_t)c * lda] = S[r * DLD + c];
            else if (r == c) A[r + (int64_t)c * lda] = ldiag[r];
            Dinv[r + (int64_t)c * ldd] = (r >= c) ? S[c * DLD + r] : 0.0;
        }
    }
    __syncthreads();
    DIAG_MARK(9);
}

// Right-looking Cholesky with one-panel look-ahead on two streams.
//   main stream (critical path):  diag(k) -> panel(k) -> [wait bulk(k-1)] -> crit(k) -> diag(k+1) ...
//   aux stream  (bulk work)     :  [wait panel(k)] -> bulk(k)
// crit(k) is the rank-128 update of block column k+1 only, bulk(k) that of the remaining trailing matrix (columns
// >= k+2); the serial diagonal-block kernels therefore overlap the large symmetric updates.
int potrf_lower_rl(cudaStream_t stream, double* A, int64_t lda, int n, double* Dinv, int64_t ldd, int* d_status,
                   CholLookahead* la) {
    if (n <= 0) return GPIRT_B200_OK;
    constexpr size_t smem = (size_t)(DB * DLD + 3 * DB) * sizeof(double);
    static bool configured[64] = {false};
    if (first_use_on_device(configured))
        GP_CUDA(cudaFuncSetAttribute(k_diag128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
        static long long* dbg = nullptr;
        static bool dbg_on = getenv("GPIRT_DIAG_DEBUG") != nullptr;
        if (dbg_on && !dbg) cudaMalloc((void**)&dbg, 32 * sizeof(long long));
        GP_LAUNCH(k_diag128, 1, DTHREADS, smem, stream, Akk, lda, nb, Dinv + k0, ldd, d_status, dbg_on ? dbg : nullptr);
        GP_CUDA(cudaGetLastError());
        if (dbg_on && k == 1) {
            long long h[32];
            cudaStreamSynchronize(stream);
            cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "k_diag128 phases (cycles): load %lld | diag32 %lld | panel+syrk32 %lld | diag32+inv32 %lld | panel64 %lld | syrk64 %lld | f_i_64 %lld | inv64 %lld | store %lld | total %lld\n",
                    h[1]-h[0], h[2]-h[1], h[3]-h[2], h[4]-h[3], h[5]-h[4], h[6]-h[5], h[7]-h[6], h[8]-h[7], h[9]-h[8], h[9]-h[0]);
            fprintf(stderr, "step c=8 (warp 2): pivot+rcp %lld | update %lld | barrier %lld\n", h[11]-h[10], h[12]-h[11], h[13]-h[12]);
        }
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
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n || c >= n) return;
    if (r / CHOL_NB == c / CHOL_NB) X[r + (int64_t)c * ldx] = Dinv[r + (int64_t)(c % CHOL_NB) * ldd];
}

int trtri_lower(cudaStream_t stream, const double* L, int64_t ldl, int n, const double* Dinv, int64_t ldd, double* X,
                int64_t ldx, double* T, int64_t ldt) {
    if (n <= 0) return GPIRT_B200_OK;
    GP_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * n * sizeof(double), stream));
    {
        dim3 grid((unsigned)ceil_div(n, 128), (unsigned)n);
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
