// FP64 GEMM on the Blackwell FP64 tensor pipe (DMMA.8x8x4 via mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind).
//
//   C[M x N] = alpha * op(A)[M x K] * op(B)[K x N] + beta * C          (column-major everywhere)
//
// One kernel serves every dense product on the hot path:
//   nu = L Z                       (draw-f.cpp:26 / mvnormal.h:10, all m items at once)       NN, tri = A lower
//   blocked triangular solves      (draw-fstar.cpp:7,19)                                        NN / TN updates
//   mean = (S^-1 K*)^T F           (draw-fstar.cpp:25)                                          TN
//   logP^T = 1/2 F* Y^T - D |Y|^T  (draw-theta.cpp:18 in contraction form)                      NT
//   Cholesky trailing updates      (gpirtMCMC.cpp:17,78,97)                                     NT, tri = C lower
//
// Tiling: CTA tile BM x BN, k-tile 16, 4-stage cp.async ring in shared memory; each warp owns a
// (BM/WARPS_M) x (BN/WARPS_N) sub-tile as 8x8 DMMA accumulator fragments held in registers.
// Shared tiles are padded by 4 doubles so that the 8x4 / 4x8 fragment gathers (lane = 4*g + t reads (g, t)) touch
// 16 distinct 8-byte words per half-warp: conflict-free for both M/N-contiguous and K-contiguous operands.
// Loads are 8-byte cp.async with zero-fill predication, so any leading dimension / offset / ragged edge is legal.
#pragma once
#include "common.cuh"

namespace gpirt {

enum GemmTri : int {
    TRI_NONE = 0,
    TRI_A_LOWER = 1,  // op(A)(m,k) == 0 for k > m : k-tiles beyond the row block are skipped
    TRI_C_LOWER = 2,  // only C(row >= col) is written (symmetric rank-k update of a lower-stored matrix)
    TRI_A_UPPER = 3,  // op(A)(m,k) == 0 for k < m : k-tiles before the row block are skipped
    TRI_B_LOWER = 4   // op(B)(k,n) == 0 for k < n : k-tiles before the column block are skipped
};

struct GemmArgs {
    int M = 0, N = 0, K = 0;
    const double* A = nullptr; int64_t lda = 0;
    const double* B = nullptr; int64_t ldb = 0;
    double* C = nullptr; int64_t ldc = 0;
    double alpha = 1.0, beta = 0.0;
    int tri = TRI_NONE;
    int b_abs = 0;      // use |op(B)| (observed-mask operand of the theta contraction)
    int batch = 1;      // blockIdx.z indexes independent problems at fixed pointer strides
    int64_t strideA = 0, strideB = 0, strideC = 0;
    int force_big = 0;  // always use the 128 x 128 tile (needed when C aliases A or B and N <= 128: one N tile)
    // split-K for thin products (few output tiles, long K): blockIdx.y owns a contiguous range of k-tiles and writes its
    // partial tile to `ws`; the CTA that arrives last at the tile's counter adds the partials IN SPLIT ORDER (the result
    // does not depend on the arrival order: deterministic) and applies alpha / beta.  ws: tiles x splitk x BM x BN doubles;
    // ws_count: one int per tile (x batch), zero before the first use — the finishing CTA resets it.
    int splitk = 1;
    double* ws = nullptr;
    int* ws_count = nullptr;
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_PAD = 4;
constexpr int GEMM_GROUP_N = 8;

__device__ __forceinline__ void cp_async_f64(double* smem_dst, const double* gmem_src, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = pred ? 8 : 0;  // src-size 0 => destination zero-filled, source not read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int BM, int BN>
constexpr size_t gemm_smem_bytes() {
    return (size_t)GEMM_STAGES * (size_t)(BM + BN) * (GEMM_BK + GEMM_PAD) * sizeof(double);
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool TA, bool TB, bool BABS>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, 1) gemm_f64_kernel(const GemmArgs g) {
    constexpr int THREADS = WARPS_M * WARPS_N * 32;
    constexpr int BK = GEMM_BK, PAD = GEMM_PAD, STAGES = GEMM_STAGES;
    constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;  // warp tile
    constexpr int MI = WTM / 8, NI = WTN / 8;
    constexpr int A_STAGE = BM * (BK + PAD), B_STAGE = BN * (BK + PAD);  // doubles (upper bound for both layouts)
    constexpr int LDA_S = TA ? (BK + PAD) : (BM + PAD);  // TA: As[m][k] ; else As[k][m]
    constexpr int LDB_S = TB ? (BN + PAD) : (BK + PAD);  // TB: Bs[k][n] ; else Bs[n][k]
    static_assert((BK * BM) % THREADS == 0 && (BK * BN) % THREADS == 0, "tile/threads mismatch");
    static_assert(BK * (BM + PAD) <= A_STAGE && BK * (BN + PAD) <= B_STAGE, "stage size");

    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    // Tile order (1-D grid per batch entry): column tiles are taken in groups of GEMM_GROUP_N; inside a group the CTAs
    // walk the row tiles (heaviest first when op(A) is lower triangular) with the group's column tiles innermost.  A
    // group's B panel (GROUP_N x 128 columns) then stays in L2 while every row tile consumes it, instead of the whole
    // B matrix being streamed from HBM once per row tile.
    const int mtiles = (g.M + BM - 1) / BM, ntiles = (g.N + BN - 1) / BN;
    int mt, nt;
    {
        const int lin = (int)blockIdx.x, per_group = GEMM_GROUP_N * mtiles;
        const int grp = lin / per_group, rem = lin % per_group;
        const int gw = min(GEMM_GROUP_N, ntiles - grp * GEMM_GROUP_N);   // width of this (possibly last, narrower) group
        mt = rem / gw;
        nt = grp * GEMM_GROUP_N + rem % gw;
        if (g.tri == TRI_A_LOWER) mt = mtiles - 1 - mt;
    }
    const int m0 = mt * BM, n0 = nt * BN;
    if (g.tri == TRI_C_LOWER && n0 > m0 + BM - 1) return;  // tile entirely above the diagonal
    const double* __restrict__ gA = g.A + (int64_t)blockIdx.z * g.strideA;
    const double* __restrict__ gB = g.B + (int64_t)blockIdx.z * g.strideB;
    double* __restrict__ gC = g.C + (int64_t)blockIdx.z * g.strideC;

    int k_begin = 0, k_end = g.K;
    if (g.tri == TRI_A_LOWER) k_end = min(g.K, m0 + BM);
    if (g.tri == TRI_A_UPPER) k_begin = (m0 / BK) * BK;
    if (g.tri == TRI_B_LOWER) k_begin = (n0 / BK) * BK;
    int nkt = (k_end - k_begin + BK - 1) / BK;
    int kt0 = 0;                                  // first k-tile of this CTA's share
    if (g.splitk > 1) {
        const int lo = (int)((int64_t)nkt * blockIdx.y / g.splitk), hi = (int)((int64_t)nkt * (blockIdx.y + 1) / g.splitk);
        kt0 = lo; nkt = hi - lo;
    }

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int wm0 = (warp % WARPS_M) * WTM, wn0 = (warp / WARPS_M) * WTN;

    // ---- loader state: every thread owns RA (RB) fixed (m,k) slots of the A (B) tile; only the k offset moves from
    // one k-tile to the next, so the global pointers are advanced by a constant and the shared offsets are constants.
    constexpr int RA = (BK * BM) / THREADS, RB = (BK * BN) / THREADS;
    constexpr int A_SLOT = TA ? THREADS / BK : THREADS / BM;   // TA: rows of m per pass ; else rows of k per pass
    constexpr int B_SLOT = TB ? THREADS / BN : THREADS / BK;   // TB: rows of k per pass ; else rows of n per pass
    const int a_fast = TA ? tid % BK : tid % BM, a_slow = TA ? tid / BK : tid / BM;   // fast = contiguous index
    const int b_fast = TB ? tid % BN : tid % BK, b_slow = TB ? tid / BN : tid / BK;
    // TA: fast = k, slow = m.  !TA: fast = m, slow = k.   TB: fast = n, slow = k.  !TB: fast = k, slow = n.
    const double* pa = TA ? gA + (int64_t)(k_begin + a_fast) + (int64_t)(m0 + a_slow) * g.lda
                          : gA + (int64_t)(m0 + a_fast) + (int64_t)(k_begin + a_slow) * g.lda;
    const double* pb = TB ? gB + (int64_t)(n0 + b_fast) + (int64_t)(k_begin + b_slow) * g.ldb
                          : gB + (int64_t)(k_begin + b_fast) + (int64_t)(n0 + b_slow) * g.ldb;
    const int64_t a_pass = (int64_t)A_SLOT * g.lda, b_pass = (int64_t)B_SLOT * g.ldb;      // pointer step between passes
    const int64_t a_tile = TA ? (int64_t)BK : (int64_t)BK * g.lda;                         // pointer step between k-tiles
    const int64_t b_tile = TB ? (int64_t)BK * g.ldb : (int64_t)BK;
    const int a_soff = TA ? a_slow * LDA_S + a_fast : a_slow * LDA_S + a_fast;
    const int b_soff = TB ? b_slow * LDB_S + b_fast : b_slow * LDB_S + b_fast;
    // (TA: As[m][k] -> slow*LDA_S + fast ; !TA: As[k][m] -> slow*LDA_S + fast : same form, slow indexes the padded row)
    const bool a_fix_ok = TA ? true : (m0 + a_fast < g.M);     // row validity that does not depend on the pass
    const bool b_fix_ok = TB ? (n0 + b_fast < g.N) : true;

    // issue passes [r0, r1) of the A and B loads of k-tile kt into `stage`
    auto load_part = [&](int stage, int kt, int part, int nparts) {
        const int k0 = k_begin + kt * BK;
        double* as = As + stage * A_STAGE + a_soff;
        double* bs = Bs + stage * B_STAGE + b_soff;
        const double* ga = pa + (int64_t)kt * a_tile;
        const double* gb = pb + (int64_t)kt * b_tile;
#pragma unroll
        for (int r = 0; r < RA; ++r) {
            if (r * nparts / RA != part) continue;
            bool ok;
            if (TA) ok = (k0 + a_fast < k_end) && (m0 + a_slow + r * A_SLOT < g.M);
            else ok = a_fix_ok && (k0 + a_slow + r * A_SLOT < k_end);
            cp_async_f64(as + r * A_SLOT * LDA_S, ok ? ga + r * a_pass : gA, ok);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (r * nparts / RB != part) continue;
            bool ok;
            if (TB) ok = b_fix_ok && (k0 + b_slow + r * B_SLOT < k_end);
            else ok = (k0 + b_fast < k_end) && (n0 + b_slow + r * B_SLOT < g.N);
            cp_async_f64(bs + r * B_SLOT * LDB_S, ok ? gb + r * b_pass : gB, ok);
        }
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) load_part(s, kt0 + s, 0, 1);
        cp_async_commit();
    }
    constexpr int KSTEPS = BK / 4;
    for (int it = 0; it < nkt; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        const bool do_load = nxt < nkt;
        const int lstage = nxt % STAGES;
        const double* as = As + (it % STAGES) * A_STAGE;
        const double* bs = Bs + (it % STAGES) * B_STAGE;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int kk = ks * 4;
            if (do_load) load_part(lstage, kt0 + nxt, ks, KSTEPS);   // the next tile's loads are spread over the k-steps
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int m = wm0 + i * 8 + gq, k = kk + tq;
                a[i] = TA ? as[m * LDA_S + k] : as[k * LDA_S + m];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int n = wn0 + j * 8 + gq, k = kk + tq;
                double w = TB ? bs[k * LDB_S + n] : bs[n * LDB_S + k];
                b[j] = BABS ? fabs(w) : w;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    if (g.splitk > 1) {   // partial tile -> workspace; the last CTA of the tile reduces in split order
        __shared__ int s_last;
        const int tile_lin = (int)(blockIdx.z * gridDim.x + blockIdx.x);
        double* wsp = g.ws + ((int64_t)tile_lin * g.splitk + blockIdx.y) * (BM * BN) + tid;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                __stcg(wsp + ((i * NI + j) * 2 + 0) * THREADS, acc[i][j][0]);
                __stcg(wsp + ((i * NI + j) * 2 + 1) * THREADS, acc[i][j][1]);
            }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int prev = atomicAdd(g.ws_count + tile_lin, 1);
            s_last = (prev == g.splitk - 1);
            if (s_last) g.ws_count[tile_lin] = 0;   // ready for the next launch on this stream
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const double* rd = g.ws + (int64_t)tile_lin * g.splitk * (BM * BN) + tid;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int sp = 0; sp < g.splitk; ++sp, rd += BM * BN) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    acc[i][j][0] += __ldcg(rd + ((i * NI + j) * 2 + 0) * THREADS);
                    acc[i][j][1] += __ldcg(rd + ((i * NI + j) * 2 + 1) * THREADS);
                }
        }
    }

    // epilogue: C fragment (row = g, cols = 2t, 2t+1).  Two passes so that, with beta != 0, all loads of the old C are
    // in flight together instead of being serialised behind the stores (the compiler cannot prove they do not alias).
    const bool has_beta = (g.beta != 0.0);
    if (has_beta) {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + wm0 + i * 8 + gq;
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = n0 + wn0 + j * 8 + 2 * tq + h;
                    double old = 0.0;
                    if (row < g.M && col < g.N && !(g.tri == TRI_C_LOWER && col > row))
                        old = __ldcg(gC + (int64_t)row + (int64_t)col * g.ldc);
                    acc[i][j][h] = g.alpha * acc[i][j][h] + g.beta * old;
                }
        }
    } else {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) { acc[i][j][0] *= g.alpha; acc[i][j][1] *= g.alpha; }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm0 + i * 8 + gq;
        if (row >= g.M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + wn0 + j * 8 + 2 * tq + h;
                if (col >= g.N) continue;
                if (g.tri == TRI_C_LOWER && col > row) continue;
                gC[(int64_t)row + (int64_t)col * g.ldc] = acc[i][j][h];
            }
        }
    }
}

// host side (gemm.cu)
int gemm_f64(cudaStream_t stream, bool ta, bool tb, const GemmArgs& g);

}  // namespace gpirt
