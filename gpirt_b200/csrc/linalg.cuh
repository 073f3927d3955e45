// Blocked FP64 factorisation / triangular solves built on the DMMA GEMM (gemm_f64.cuh).
#pragma once
#include <functional>
#include <vector>

#include "common.cuh"

namespace gpirt {

constexpr int DIAG_NB = 64;  // diagonal blocks inverted inside one CTA by the stand-alone triangular solve (gpirt_b200_trsm_lower)

// Dinv <- inverses of the 64 x 64 diagonal blocks of an existing lower-triangular L (one CTA per block).
int trtri_diag_blocks(cudaStream_t stream, const double* L, int64_t ldl, int n, double* Dinv, int64_t ldd);

// B <- L^-1 B   (trans = false)   or   B <- L^-T B   (trans = true);  L n x n lower, B n x nrhs, in place.
int trsm_left_lower(cudaStream_t stream, bool trans, int n, int nrhs, const double* L, int64_t ldl,
                    const double* Dinv, int64_t ldd, double* B, int64_t ldb);

constexpr int CHOL_NB = 128;  // panel width of the right-looking factorisation; diagonal blocks done by one CTA

// Right-looking blocked Cholesky (lower), panel width 128.  Per panel: k_diag128 (chol_diag.cuh) factorises AND inverts
// the 128 x 128 diagonal block in one CTA; k_panel_update (chol_panel.cuh) multiplies the panel below by that inverse and
// applies its rank-128 update to the next block column; the rest of the trailing matrix is updated on the DMMA GEMM.
// Dinv128 (n x 128, ld ldd, 16-byte aligned, ldd even) receives the inverse of every diagonal block of L; only the lower
// triangles are written, so the caller zero-initialises Dinv128 once.  d_flags: scratch, ceil(n / 128) ints.
struct CholLookahead {      // second stream + events for the one-panel look-ahead (owned by the caller)
    cudaStream_t aux = nullptr;
    std::vector<cudaEvent_t> ev_panel, ev_bulk;
    // optional hook: called on the host right after block column k of L has been enqueued as final (event `done`,
    // recorded on the main stream) — lets the caller start work that consumes finished columns while the chain runs
    std::function<int(int k, int nblk, cudaEvent_t done)> after_panel;
};
int potrf_lower_rl(cudaStream_t stream, double* A, int64_t lda, int n, double* Dinv128, int64_t ldd, int* d_status,
                   int* d_flags, CholLookahead* la = nullptr);

// X = L^-1 (n x n, lower; strict upper zero) by batched recursive doubling from the 128 x 128 block inverses:
//   inv([[L11,0],[L21,L22]]) = [[X11,0],[-X22 L21 X11, X22]], all pairs of one level in a single batched GEMM launch.
// T is n x n scratch.  max_block > 0 (a power-of-two multiple of 128) stops the doubling there: X then holds the inverses of
// the max_block x max_block diagonal blocks of L only (zeros elsewhere) — the leaf blocks of a coarser substitution.
// splitk_ws / splitk_count (optional): workspace of the deterministic split-K of gemm_f64 — products with K >= 512 are
// split min(8, K / 128) ways (the late levels are a few tiles with a long K).
int trtri_lower(cudaStream_t stream, const double* L, int64_t ldl, int n, const double* Dinv128, int64_t ldd, double* X,
                int64_t ldx, double* T, int64_t ldt, int max_block = 0, double* splitk_ws = nullptr, int* splitk_count = nullptr);

// The same block inverses built PROGRESSIVELY behind a right-looking factorisation: call once per finished panel k
// (k = 0 .. nblk-1 in order, on one stream).  Step k copies the 128-block inverse of panel k into X and merges every pair
// of diagonal blocks of order 128, 256, ... < max_block that panel k completes (binary-counter pattern), so that after
// the last panel only the merges that involve it remain: one pair per level.  X must be zero outside the written blocks
// (zero it once; the written pattern is the same for every factorisation of the same order).
int trtri_lower_step(cudaStream_t stream, const double* L, int64_t ldl, int n, const double* Dinv128, int64_t ldd, double* X,
                     int64_t ldx, double* T, int64_t ldt, int max_block, int k, double* splitk_ws, int* splitk_count);

}  // namespace gpirt
