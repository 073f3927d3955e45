// Blocked FP64 factorisation / triangular solves built on the DMMA GEMM (gemm_f64.cuh).
#pragma once
#include "common.cuh"

namespace gpirt {

constexpr int DIAG_NB = 64;  // diagonal blocks factorised + inverted inside one CTA

// In-place lower Cholesky of the n x n matrix A (only the lower triangle is read or written; whatever the caller
// left in the strict upper triangle stays).  Dinv (n x 64, leading dimension ldd) receives the inverse of every
// 64 x 64 diagonal block of L (block b at rows 64b..), which the triangular solves below consume.
// d_status (device int) is set non-zero if a pivot is not positive (chol(): decomposition failed).
int potrf_lower(cudaStream_t stream, double* A, int64_t lda, int n, double* Dinv, int64_t ldd, int* d_status);

// Dinv <- inverses of the 64 x 64 diagonal blocks of an existing lower-triangular L (one CTA per block).
int trtri_diag_blocks(cudaStream_t stream, const double* L, int64_t ldl, int n, double* Dinv, int64_t ldd);

// B <- L^-1 B   (trans = false)   or   B <- L^-T B   (trans = true);  L n x n lower, B n x nrhs, in place.
int trsm_left_lower(cudaStream_t stream, bool trans, int n, int nrhs, const double* L, int64_t ldl,
                    const double* Dinv, int64_t ldd, double* B, int64_t ldb);

// X <- X L^-T  (solve X L^T = B in place; X is rows x n), used by the Cholesky panel step.
int trsm_right_lower_t(cudaStream_t stream, int rows, int n, const double* L, int64_t ldl, const double* Dinv,
                       int64_t ldd, double* X, int64_t ldx);

}  // namespace gpirt
