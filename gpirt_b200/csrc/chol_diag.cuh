// 128 x 128 diagonal block of the right-looking Cholesky (arma::chol -> dpotrf, reference src/gpirtMCMC.cpp:17,78,97):
// Cholesky factor AND its inverse in one CTA of 8 warps.  This kernel is the serial link of the factorisation chain
// (one launch per 128 columns, nothing else can start before it ends), so it is written for latency, not throughput.
// A warp issues in order, so whatever sits in the instruction stream of the warp that carries the pivot chain delays the
// next pivot; the design keeps that stream minimal and gives everything else to other warps:
//
//   * the lower triangle arrives in 16-byte pieces with all loads of a thread in flight at once (one L2 round trip; 128
//     per-column bulk-async copies were tried and cost ~60 cycles each in the copy engine, 8k cycles in all), and every
//     finished 32-column panel of L is written back by otherwise idle warps while the next panel is factorised;
//   * the block lives in shared memory COLUMN-major with leading dimension 132 (= 4 mod 16): lane-per-row register
//     loads, DMMA fragment gathers in both orientations (bank 4g+t and 4t+g) and the global copies are conflict-free;
//   * right-looking with 32-column panels.  The 32 x 32 diagonal block of a panel is eliminated by ONE leader warp in
//     registers (lane = row, its 32 entries in 32 registers; the pivot column is handed over through shared memory with
//     one __syncwarp): no block barrier on the 128-step dependency chain
//         pivot -> reciprocal -> next pivot      (STS/LDS hand-over + MUFU.RCP64H + 2 Newton steps + 1 DFMA per column).
//     The leader publishes every finished column (unscaled entries, pivot, 1/pivot) and bumps a sequence flag;
//   * the rows below the diagonal block are eliminated by follower warps (lane = row) that trail the leader by about one
//     column, and a scaler warp turns the published columns into L (the 1/sqrt(pivot) scaling never enters the chain);
//   * rank-32 trailing updates and the off-diagonal blocks of the inverse run on the FP64 tensor pipe (DMMA.8x8x4)
//     straight out of shared memory, 4 to 8 accumulator chains per warp (a dependent DMMA issues every 26 cycles, an
//     independent one every 16); the 32 x 32 diagonal blocks of the inverse by forward substitution, lane = column;
//   * every piece of the inverse goes to global memory as soon as it is final (off-diagonal blocks straight from the
//     DMMA accumulator fragments), so only the last 64 x 64 block is stored after the last product.
//
// Storage while the kernel runs: strictly below the diagonal L, on the diagonal 1/L_rr (= the inverse's diagonal; L_rr
// itself goes to ldiag[]), strictly above the diagonal the TRANSPOSE of X = L^-1 (X(r,c), r > c, lives at (c, r)).
// Entries above the diagonal that nothing has written yet are uninitialised; every read of them is masked.
#pragma once
#include "gemm_f64.cuh"
#include "linalg.cuh"

namespace gpirt {
namespace diag {

constexpr int DB = CHOL_NB;          // 128
constexpr int DLD = DB + 4;          // 132
constexpr int NBK = 32;              // inner panel width
constexpr int NPAN = DB / NBK;       // 4
constexpr int DWARPS = 8, DTHREADS = DWARPS * 32;
constexpr int SCALER_WARP = DWARPS - 1;   // warp 4 shares its scheduler with the leader and stays idle during the panels

struct __align__(16) Smem {
    double S[DB * DLD];              // element (r, c) at S[c * DLD + r]
    double2 pairbuf[(NBK / 2) * NBK];   // columns (c, c+1) of the current panel's diagonal block, unscaled: pairbuf[pair * 32 + row]
    double pbuf[(NBK / 2) * 8];         // per pair: p, p' = r - q s, 1/p, s = q/p, 1/p'  (stride 8)
    double ldiag[DB];                // L_rr
    double rsbuf[DWARPS][NBK];       // per-warp 1 / sqrt(pivot) of the current panel's columns
    int flag;                        // number of panel columns published so far (monotone over the four panels)
};

__device__ __forceinline__ double& at(double* S, int r, int c) { return S[c * DLD + r]; }

// fast 1/p for p > 0: hardware reciprocal seed (MUFU.RCP64H, ~20 bits) + two Newton steps (within 1 ulp)
__device__ __forceinline__ double fast_rcp(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    r = fma(r, fma(-p, r, 1.0), r);
    r = fma(r, fma(-p, r, 1.0), r);
    return r;
}
// fast 1/sqrt(p) for p > 0: MUFU.RSQ64H seed + two Newton steps (never on the dependency chain)
__device__ __forceinline__ double fast_rsqrt(double p) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    y = fma(0.5 * y, fma(-(p * y), y, 1.0), y);
    y = fma(0.5 * y, fma(-(p * y), y, 1.0), y);
    return y;
}
// Sequence flag in shared memory.  Shared-memory accesses of one SM are performed in issue order, so a volatile store
// after the data stores of the same thread (and after a __syncwarp for the other lanes' stores) publishes them; an
// acquire/release pair costs ~200 cycles per hand-over here (tools/lat_fp64.cu) and would sit on the pivot chain.
__device__ __forceinline__ void flag_store(int* p, int v) {
    asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_load(const int* p) {
    int v;
    asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
// wait until the flag reaches `target`; a waiting warp backs off so that its polling does not compete with the leader's
// shared-memory hand-over (measured: three spinning followers slowed the leader by ~15%)
__device__ __forceinline__ int flag_wait(const int* p, int target, unsigned backoff_ns) {
    int v = flag_load(p);
    while (v < target) { __nanosleep(backoff_ns); v = flag_load(p); }
    return v;
}
// Everything a consumer reads after the flag must be a volatile access as well: ptxas keeps volatile accesses in program
// order among themselves but hoists ordinary shared-memory loads above the polling loop (seen in the SASS: the column
// was read before the flag said it was there).
__device__ __forceinline__ double ldv(const double* p) {
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ double2 ldv2(const double* p) {
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void stv(double* p, double v) {
    asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v) : "memory");
}

// ---- one 32-column panel, eliminated two columns at a time (2 x 2 pivot blocks) -------------------------------------
// Symmetric block elimination of columns (c, c+1) with D = [[p, q], [q, r]] (p = a_cc, q = a_{c+1,c}, r = a_{c+1,c+1}):
//     a_ij -= [a_ic  a_i,c+1] D^-1 [a_jc  a_j,c+1]^T        (j > c+1)
// which is exactly two consecutive scalar steps of the symmetric elimination (the second pivot is p' = r - q^2/p =
// det(D) / p).  One pair step costs ONE reciprocal on the dependency chain (1 / det; 1 / p runs beside it) instead of two
// dependent ones, and one shared-memory hand-over instead of two: 16 links per panel instead of 32.
//   multipliers of row i:   m = a_ic / p,   b' = a_i,c+1 - a_ic s  (s = q / p: the updated column c+1),   m' = b' / p',
//                           u = m - m' s,   v = m'         so that     a_ij -= u a_jc + v a_j,c+1
//   L:   l_ic = a_ic / sqrt(p),   l_i,c+1 = b' / sqrt(p').
// The entries of the NEXT pair's columns are updated in the form  a_ij - w_ij / det  with
//     w_ij = (a_ic r - a_i,c+1 q) a_jc + (a_i,c+1 p - a_ic q) a_j,c+1
// so that everything but the final DFMA is computed while the reciprocal of det is in flight.
// pairbuf[pair][row] = (a_row,c , a_row,c+1) unscaled;  pbuf[pair] = (p, p', 1/p, s, 1/p').
constexpr int NPAIR = NBK / 2, PB_STRIDE = 8;

// Leader warp (lane = row of the diagonal block, its 32 panel entries in registers).  The whole routine is one basic
// block, so ptxas can fill the chain's stall slots with the trailing updates.  (Measured alternatives: a leader that
// carries only the chain, with a helper warp owning the later columns of the same rows and handing each pair back through
// shared memory, was 2x SLOWER — the two flag hand-overs per pair cost more than the trailing updates they remove.)
__device__ __noinline__ void panel32_lead(Smem& sm, int o, int seq0, int* status) {
    double* S = sm.S;
    const int lane = threadIdx.x & 31;
    double r[NBK];
#pragma unroll
    for (int j = 0; j < NBK; ++j) r[j] = (j <= lane) ? at(S, o + lane, o + j) : 0.0;
    bool bad = false;
#pragma unroll
    for (int pr = 0; pr < NPAIR; ++pr) {
        const int c = 2 * pr;
        const double a = r[c], b = r[c + 1];
        double2* buf = sm.pairbuf + pr * NBK;
        buf[lane] = make_double2(a, b);
        __syncwarp();
        const double2 d0 = buf[c], d1 = buf[c + 1];
        const double p = d0.x, q = d1.x, rr = d1.y;
        const double det = fma(p, rr, -(q * q));
        const double dinv = fast_rcp(det);
        const double pinv = fast_rcp(p);
        const double e0 = fma(a, rr, -(b * q)), e1 = fma(b, p, -(a * q));     // (u, v) = (e0, e1) / det
        if (c + 2 < NBK) {   // the next pair's two columns first: they carry the chain
            const double2 n0 = buf[c + 2], n1 = buf[c + 3];
            r[c + 2] = fma(-fma(e1, n0.y, e0 * n0.x), dinv, r[c + 2]);
            r[c + 3] = fma(-fma(e1, n1.y, e0 * n1.x), dinv, r[c + 3]);
        }
        const double s = q * pinv, p2 = fma(-q, s, rr), pinv2 = p * dinv;
        if (lane == 0) {
            double* pb = sm.pbuf + pr * PB_STRIDE;
            stv(pb, p); stv(pb + 1, p2); stv(pb + 2, pinv); stv(pb + 3, s); stv(pb + 4, pinv2);
            flag_store(&sm.flag, seq0 + c + 2);
        }
        bad |= !(p > 0.0) | !(det > 0.0);
        const double u = e0 * dinv, v = e1 * dinv;
#pragma unroll
        for (int j = c + 4; j < NBK; ++j) {
            const double2 aj = buf[j];
            r[j] = fma(-v, aj.y, fma(-u, aj.x, r[j]));
        }
    }
    if (bad && lane == 0) atomicExch(status, 1);   // not positive definite (or NaN): chol(): decomposition failed
}

// Follower warp: the same pair steps for 32 rows below the diagonal block, two published pairs at a time (one poll and
// one batch of loads per four columns keeps a follower ahead of the leader's pace; the panel ends one short batch after
// the leader).  Its rows of L are scaled by 1/sqrt(pivot) and written once at the end.
constexpr int FOLLOW_PAIRS = 2;
__device__ __noinline__ void panel32_follow(Smem& sm, int o, int row, int seq0, int warp) {
    double* S = sm.S;
    const int lane = threadIdx.x & 31;
    double r[NBK];
#pragma unroll
    for (int j = 0; j < NBK; ++j) r[j] = at(S, row, o + j);
    int have = 0;
#pragma unroll
    for (int pb0 = 0; pb0 < NPAIR; pb0 += FOLLOW_PAIRS) {
        if (have < seq0 + 2 * (pb0 + FOLLOW_PAIRS)) have = flag_wait(&sm.flag, seq0 + 2 * (pb0 + FOLLOW_PAIRS), 100);
#pragma unroll
        for (int pr = pb0; pr < pb0 + FOLLOW_PAIRS; ++pr) {
            const int c = 2 * pr;
            const double* pb = sm.pbuf + pr * PB_STRIDE;
            const double2 ps = ldv2(pb + 2);                     // (1/p, s)
            const double pinv2 = ldv(pb + 4);
            const double a = r[c];
            const double bp = fma(-a, ps.y, r[c + 1]);           // b' = a_i,c+1 - a_ic s
            r[c + 1] = bp;                                       // kept for the final scaling
            const double v = bp * pinv2, u = fma(-v, ps.y, a * ps.x);
            const double* buf = reinterpret_cast<const double*>(sm.pairbuf + pr * NBK);
#pragma unroll
            for (int j = c + 2; j < NBK; ++j) {
                const double2 aj = ldv2(buf + 2 * j);
                r[j] = fma(-v, aj.y, fma(-u, aj.x, r[j]));
            }
        }
    }
    double* rs = sm.rsbuf[warp];                                 // lane c: 1 / sqrt(pivot of column c)
    rs[lane] = fast_rsqrt(ldv(sm.pbuf + (lane >> 1) * PB_STRIDE + (lane & 1)));
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NBK; ++c) at(S, row, o + c) = r[c] * rs[c];
}

// Scaler warp: once the leader has published all 16 pairs, L entries of the diagonal block itself
// (l_ic = a_ic / sqrt(p), l_i,c+1 = (a_i,c+1 - a_ic s) / sqrt(p')), its diagonal slots (1 / L_cc) and L_cc.
__device__ __forceinline__ void panel32_scale(Smem& sm, int o, int seq0, int warp) {
    double* S = sm.S;
    const int lane = threadIdx.x & 31;
    flag_wait(&sm.flag, seq0 + NBK, 300);
    double* rs = sm.rsbuf[warp];
    const double p_mine = ldv(sm.pbuf + (lane >> 1) * PB_STRIDE + (lane & 1));
    const double rs_mine = fast_rsqrt(p_mine);
    rs[lane] = rs_mine;
    at(S, o + lane, o + lane) = rs_mine;
    sm.ldiag[o + lane] = p_mine * rs_mine;
    __syncwarp();
#pragma unroll 1
    for (int p0 = 0; p0 < NPAIR; p0 += 4) {   // four pairs of volatile loads in flight, then their stores
        double2 ab[4];
        double sv[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            ab[w] = ldv2(reinterpret_cast<const double*>(sm.pairbuf + (p0 + w) * NBK + lane));
            sv[w] = ldv(sm.pbuf + (p0 + w) * PB_STRIDE + 3);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int c = 2 * (p0 + w);
            if (lane > c) at(S, o + lane, o + c) = ab[w].x * rs[c];
            if (lane > c + 1) at(S, o + lane, o + c + 1) = fma(-ab[w].x, sv[w], ab[w].y) * rs[c + 1];
        }
    }
}

// Columns [o, o+32) of L are final after their panel: idle warps write them to global memory while the next panel runs
// (part = 0..parts-1 splits the 32 columns), so that only the inverse is left to store when the kernel ends.
__device__ __forceinline__ void store_L_columns(Smem& sm, double* __restrict__ A, int64_t lda, int nb, int o, int part, int parts) {
    double* S = sm.S;
    const int lane = threadIdx.x & 31;
    for (int c = o + part; c < o + NBK; c += parts) {
        if (c >= nb) break;
        for (int r = c + lane; r < nb; r += 32) A[r + (int64_t)c * lda] = (r == c) ? sm.ldiag[c] : at(S, r, c);
    }
}

// 32 x 32 diagonal block of the inverse, lane = column j:  x_j = L_bb^-1 e_j  by column-oriented forward substitution
// (all L reads are warp-wide broadcasts); X(k, j), k > j, is stored transposed at (o + j, o + k).
__device__ __noinline__ void inv32(Smem& sm, int o, int lane) {
    double* S = sm.S;
    double acc[NBK];
#pragma unroll
    for (int i = 0; i < NBK; ++i) acc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < NBK; ++k) {
        const double* col = &at(S, o, o + k);                     // column k of L_bb (diagonal entry = 1 / L_kk)
        const double xk = acc[k] * col[k];
        if (k > lane) at(S, o + lane, o + k) = xk;
#pragma unroll
        for (int ii = (k + 1) & ~1; ii < NBK; ii += 2) {
            const double2 v = *reinterpret_cast<const double2*>(col + ii);
            if (ii >= k + 1) acc[ii] = fma(-v.x, xk, acc[ii]);
            acc[ii + 1] = fma(-v.y, xk, acc[ii + 1]);
        }
    }
}

// Rank-32 update of the trailing block after panel [o, o+32):  A22 -= L21 L21^T  (lower triangle).  Work items are
// 16 x 16 blocks = 2 x 2 DMMA fragments (four accumulator chains, one A/B gather per DMMA), dealt round-robin to the warps.
__device__ __forceinline__ void syrk32(double* S, int o, int warp, int lane) {
    const int ro = o + NBK, nb2 = (DB - ro) / 16, total = nb2 * (nb2 + 1) / 2;
    const int g = lane >> 2, t = lane & 3;
    for (int f = warp; f < total; f += DWARPS) {
        int RB = 0;
        while ((RB + 1) * (RB + 2) / 2 <= f) ++RB;
        const int CB = f - RB * (RB + 1) / 2;
        const int ri = ro + RB * 16 + g, ci = ro + CB * 16 + g;
        double acc[2][2][2] = {};
#pragma unroll
        for (int ks = 0; ks < NBK / 4; ++ks) {
            const int k = o + ks * 4 + t;
            const double a0 = at(S, ri, k), a1 = at(S, ri + 8, k), b0 = at(S, ci, k), b1 = at(S, ci + 8, k);
            dmma_8x8x4(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma_8x8x4(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma_8x8x4(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma_8x8x4(acc[1][1][0], acc[1][1][1], a1, b1);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int i = RB * 16 + u * 8 + g, j = CB * 16 + v * 8 + 2 * t;
                if (i >= j) at(S, ro + i, ro + j) -= acc[u][v][0];
                if (i >= j + 1) at(S, ro + i, ro + j + 1) -= acc[u][v][1];
            }
    }
}

// Off-diagonal block of the inverse:  X21 = -X22 (L21 X11)  for the H x H block at rows ro.., columns co.. ; H/8 warps
// cooperate (wi = 0..H/8-1).  First product: warp wi owns the 8-row block wi of T = L21 X11 (X11 lower triangular: the
// work per row block is uniform); second product: warp wi owns the 8-COLUMN block wi of X21 = -X22 T (X22 lower
// triangular: per column block the work is uniform too).  Contains block-wide barriers: every warp of the CTA calls it
// the same number of times.
template <int H, bool TO_SMEM>
__device__ __forceinline__ void inv_offdiag(double* S, int ro, int co, int wi, int lane, double* __restrict__ Dinv, int64_t ldd, int nb) {
    constexpr int NF = H / 8;
    const int g = lane >> 2, t = lane & 3;
    double acc[NF][2];
    // T(i,j) = sum_{k >= j} L21(i,k) X11(k,j),  X11(k,j) at (co+j, co+k);  T goes to the X21 slot (co+j, ro+i).
    // k is fully unrolled: the structural-zero skips are decided at compile time and acc[] stays in registers.
#pragma unroll
    for (int q = 0; q < NF; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < H; k0 += 4) {
        const double av = at(S, ro + wi * 8 + g, co + k0 + t);
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            if (k0 + 3 < q * 8) continue;                         // X11(k,j) = 0 for k < j
            const int j = q * 8 + g, k = k0 + t;
            const double x = at(S, co + j, co + k);
            dmma_8x8x4(acc[q][0], acc[q][1], av, (k0 >= q * 8 + 8 || k >= j) ? x : 0.0);
        }
    }
#pragma unroll
    for (int q = 0; q < NF; ++q) {
        at(S, co + q * 8 + 2 * t, ro + wi * 8 + g) = acc[q][0];
        at(S, co + q * 8 + 2 * t + 1, ro + wi * 8 + g) = acc[q][1];
    }
    __syncthreads();
    // X21(i,j) = - sum_{k <= i} X22(i,k) T(k,j),  X22(i,k) at (ro+k, ro+i),  T(k,j) at (co+j, ro+k);  j in column block wi
#pragma unroll
    for (int q = 0; q < NF; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < H; k0 += 4) {
        const double bv = at(S, co + wi * 8 + g, ro + k0 + t);
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            if (q * 8 + 7 < k0) continue;                         // X22(i,k) = 0 for k > i
            const int i = q * 8 + g, k = k0 + t;
            const double x = at(S, ro + k, ro + i);
            dmma_8x8x4(acc[q][0], acc[q][1], (k0 + 3 < q * 8 || k <= i) ? x : 0.0, bv);
        }
    }
    // the result goes to global memory straight from the accumulator fragments (8 rows x 64 bytes per store: full
    // sectors) and, if a later level still needs it, into the X21 slot of the shared array
#pragma unroll
    for (int q = 0; q < NF; ++q) {
        const int r = ro + q * 8 + g, c = co + wi * 8 + 2 * t;
        if (r < nb && c < nb) Dinv[r + (int64_t)c * ldd] = -acc[q][0];
        if (r < nb && c + 1 < nb) Dinv[r + (int64_t)(c + 1) * ldd] = -acc[q][1];
    }
    if (TO_SMEM) {
        __syncthreads();                                          // T fully consumed before X21 replaces it
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            at(S, co + wi * 8 + 2 * t, ro + q * 8 + g) = -acc[q][0];
            at(S, co + wi * 8 + 2 * t + 1, ro + q * 8 + g) = -acc[q][1];
        }
        __syncthreads();
    }
}

// the 32 x 32 diagonal blocks of the inverse (X(r,c) at (c, r) of the shared array) to global memory, 16 columns per warp:
// lanes run down a column, i.e. along the strided direction of the shared array (8-way conflicts on 16 loads per warp —
// cheaper than staging 4 tiles on 4 of the 8 warps)
__device__ __forceinline__ void store_X_diag_blocks(Smem& sm, double* __restrict__ Dinv, int64_t ldd, int nb, int warp, int lane) {
    double* S = sm.S;
#pragma unroll 4
    for (int u = 0; u < DB / DWARPS; ++u) {
        const int c = warp * (DB / DWARPS) + u, r = (c & ~(NBK - 1)) + lane;
        if (r >= c && r < nb) Dinv[r + (int64_t)c * ldd] = at(S, c, r);
    }
}

// A: the nb x nb diagonal block (column-major, lower triangle read; L written to the lower triangle incl. diagonal);
// Dinv: nb x nb inverse of L, lower triangle written (the strict upper triangle must hold zeros: zero-initialise Dinv
// once).  nb <= 128; missing rows/columns are padded with the identity.  PROBE: phase time stamps (clock64) into dbg[] — instantiated by tools/diag_probe.cu only.
template <bool PROBE>
__global__ void __launch_bounds__(DTHREADS, 1) k_diag128(double* __restrict__ A, int64_t lda, int nb,
                                                         double* __restrict__ Dinv, int64_t ldd, int* status,
                                                         long long* dbg) {
    extern __shared__ __align__(16) unsigned char diag_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(diag_raw);
    double* S = sm.S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int mark = 0;
    auto stamp = [&]() {   // the memory clobber keeps the clock read on its side of the surrounding barriers
        if (PROBE) {
            long long t;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
            if (tid == 0) dbg[mark] = t;
            ++mark;
        }
    };
    stamp();
    if (tid == 0) sm.flag = 0;
    {
        // lower triangle in 16-byte pieces, all loads of a thread in flight together (about one L2 round trip);
        // pieces entirely above the diagonal are skipped and stay uninitialised
        const bool vec = (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
        constexpr int PIECES = (DB / 2) * DB / DTHREADS;          // 32 row-pair x column pieces per thread
        double2 v[PIECES];
#pragma unroll
        for (int u = 0; u < PIECES; ++u) {
            const int idx = tid + u * DTHREADS, p = idx % (DB / 2), c = idx / (DB / 2), r = 2 * p;
            v[u] = make_double2((r == c) ? 1.0 : 0.0, (r + 1 == c) ? 1.0 : 0.0);   // identity padding beyond nb
            if (r + 1 >= c && c < nb) {
                if (vec && r + 1 < nb) v[u] = *reinterpret_cast<const double2*>(A + r + (int64_t)c * lda);
                else {
                    if (r < nb) v[u].x = A[r + (int64_t)c * lda];
                    if (r + 1 < nb) v[u].y = A[r + 1 + (int64_t)c * lda];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PIECES; ++u) {
            const int idx = tid + u * DTHREADS, p = idx % (DB / 2), c = idx / (DB / 2), r = 2 * p;
            if (r + 1 >= c) *reinterpret_cast<double2*>(&at(S, r, c)) = v[u];
        }
        __syncthreads();
    }
    stamp();
    // ---- factorisation: four 32-column panels ----
#pragma unroll 1
    for (int b = 0; b < NPAN; ++b) {
        const int o = b * NBK;
        if (warp == 0) panel32_lead(sm, o, o, status);
        else if (warp < NPAN - b) panel32_follow(sm, o, o + NBK * warp + lane, o, warp);
        else if (warp == SCALER_WARP) panel32_scale(sm, o, o, warp);
        else if (b > 0 && warp == 6) {                                       // under panels 1..3: the previous panel's leftovers
            inv32(sm, o - NBK, lane);                                        //   inverse of its diagonal block
            store_L_columns(sm, A, lda, nb, o - NBK, 0, 1);                  //   its columns of L to global memory
        }
        stamp();
        __syncthreads();
        stamp();
        if (b + 1 < NPAN) {
            syrk32(S, o, warp, lane);
            __syncthreads();
        }
        stamp();
    }
    // ---- inverse: the last diagonal 32-block (warp 0; the others write the last panel's columns of L meanwhile), then the
    // off-diagonal blocks of the 64- and 128-level; every piece of X goes to global memory as soon as it is final, so the
    // stores drain under the remaining products and only the last 64 x 64 block is left when the kernel ends (the strict
    // upper triangle of Dinv is never written: the caller zero-initialises it once) ----
    if (warp == 0) inv32(sm, DB - NBK, lane);                              // blocks 0..2 were inverted under the later panels
    else store_L_columns(sm, A, lda, nb, DB - NBK, warp - 1, DWARPS - 1);
    __syncthreads();
    stamp();
    store_X_diag_blocks(sm, Dinv, ldd, nb, warp, lane);
    inv_offdiag<32, true>(S, warp < 4 ? 32 : 96, warp < 4 ? 0 : 64, warp & 3, lane, Dinv, ldd, nb);
    stamp();
    inv_offdiag<64, false>(S, 64, 0, warp, lane, Dinv, ldd, nb);
    stamp();
}

constexpr size_t SMEM_BYTES = sizeof(Smem);

}  // namespace diag
}  // namespace gpirt
