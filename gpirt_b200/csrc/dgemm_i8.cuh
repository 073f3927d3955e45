// FP64-accurate matrix products on the int8 tensor cores (tcgen05.mma.kind::i8): both operands are split into balanced
// base-128 digit planes and the plane products are accumulated exactly in int32 (dgemm_i8.cu).
#pragma once
#include "common.cuh"

namespace gpirt {

// One operand of C = A B^T, held as DG_S digit planes.  Row r of the operand (a row of A or a row of B^T) is the
// 56-bit fixed-point vector  2^(e_r - 55) X[r, :],  X = sum_s d_s 128^(7 - s),  d_s in [-64, 64];  plane s is the int8
// matrix d_s, K-major with row pitch k_pad, planes stacked: element (s, r, k) at planes[(s rows_pad + r) k_pad + k].
struct DigitPlanes {
    struct Map;                      // TMA tensor map (opaque: keeps <cuda.h> out of this header)
    static constexpr int S = 8;
    int rows = 0, k = 0, box_rows = 0;
    int64_t rows_pad = 0, k_pad = 0;
    int8_t* planes = nullptr;
    double* scale = nullptr;         // 2^(e_r - 6): multiplies sum_s 128^-s (digit products)
    double* partial = nullptr;       // row-max scratch of the transposing slicer
    Map* map = nullptr;
    cudaStream_t stream_for_free = nullptr;

    // box_rows: 128 for the A operand (tile rows), 64 for the B operand (tile columns)
    int init(cudaStream_t st, int rows, int k, int box_rows);
    // operand row r = src[r ld + 0 .. k)          (the contraction index is contiguous in memory: Z, f, K* solves)
    int slice_kcontig(cudaStream_t st, const double* src, int64_t ld);
    // operand row r = src[r + kk ld], kk in [k_lo, k_hi)   (rows contiguous: the Cholesky factor);  lower: entries with
    // kk > r are taken as zero.  fixed_exp = INT_MIN: row scales from the row maxima (needs the whole row: k_lo = 0,
    // k_hi = k); otherwise every row uses 2^fixed_exp as its bound (|entries| < 2^fixed_exp), which lets a matrix that
    // is still being produced (the Cholesky factor, panel by panel) be sliced in column ranges.
    int slice_mcontig(cudaStream_t st, const double* src, int64_t ld, bool lower, int k_lo, int k_hi, int fixed_exp);
    void destroy();
};

// C[i + j ldc] (+)= sum_{k in [k_lo, k_hi)} A[i, k] B[j, k]    i < A.rows, j < B.rows
// a_tri: DG_TRI_LOWER: A[i, k] = 0 for k > i (row tiles stop at the diagonal block); DG_TRI_UPPER: A[i, k] = 0 for k < i
// (row tiles start there) — the planes must hold zeros in the skipped part.  group_cols: column tiles per L2-resident
// group (0: one group).
enum { DG_TRI_NONE = 0, DG_TRI_LOWER = 1, DG_TRI_UPPER = 2 };
int dgemm_i8(cudaStream_t st, const DigitPlanes& A, const DigitPlanes& B, double* C, int64_t ldc, int a_tri,
             int k_lo, int k_hi, bool accumulate, int group_cols, bool persistent = true);

}  // namespace gpirt
