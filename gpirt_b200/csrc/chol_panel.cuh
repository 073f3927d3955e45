// Panel step of the right-looking Cholesky, fused with the update of the next block column (reference: arma::chol ->
// dpotrf, src/gpirtMCMC.cpp:17,78,97).  After the diagonal block k has been factorised and inverted (chol_diag.cuh):
//
//   P        = A21 X11^T                         the rem x 128 panel below the diagonal block (X11 = L11^-1, lower)
//   A[:, k+1] -= P P_top^T                       rank-128 update of block column k+1 only (P_top = first 128 rows of P);
//                                                the rest of the trailing matrix is updated by the bulk GEMM on the
//                                                look-ahead stream
//
// in ONE launch on the critical path of the factorisation (before: a device-to-device copy of the panel, a GEMM with
// the block inverse and a second GEMM for the update, with their launch gaps).  Every CTA owns R = 32 (or 16) rows of
// the panel: it reads them once, overwrites them in place with P and keeps P in shared memory as the A operand of the
// update.  The update needs P_top, which the first 128 / R CTAs produce: they publish their rows, fence, and bump a
// counter; the others wait for the counter (CTAs are dispatched in index order, so the producers are always resident)
// and then pull their B operands — two 8-row slices of P_top per warp — straight from L2 into registers.
// Both products run on the FP64 tensor pipe (DMMA.8x8x4), eight accumulator chains per warp.
#pragma once
#include "gemm_f64.cuh"
#include "linalg.cuh"

namespace gpirt {
namespace panel {

constexpr int PB = CHOL_NB;          // 128
constexpr int XLD = PB + 4;          // 132 = 4 mod 16: conflict-free DMMA fragment gathers (bank 4t + g)
constexpr int PTHREADS = 256;

template <int R> struct __align__(16) Smem {
    static constexpr int ALD = R + 4;         // 36 / 20 = 4 mod 16
    double X[PB * XLD];                       // X11, element (j, k) at X[k * XLD + j]
    double A[PB * ALD];                       // this CTA's rows, k-major: (row, k) at A[k * ALD + row]; A21 first, then P
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// L: the n x n matrix being factorised (column-major, ld); k0: first row/column of diagonal block k (128 wide, final);
// Dinv: inverse of that block (128 x 128, lower, zeros above).  counter: one int, zero on entry.
// PROBE: phase time stamps (clock64) of CTA `probe_cta` into dbg[] — instantiated by tools/panel_probe.cu only.
template <int R, bool PROBE = false>
__global__ void __launch_bounds__(PTHREADS, 1) k_panel_update(double* __restrict__ L, int64_t ld, int n, int k0,
                                                              const double* __restrict__ Dinv, int64_t ldd, int* counter,
                                                              long long* dbg = nullptr, int probe_cta = 0) {
    constexpr int RF = R / 8, ALD = Smem<R>::ALD;
    extern __shared__ __align__(16) unsigned char panel_raw[];
    Smem<R>& sm = *reinterpret_cast<Smem<R>*>(panel_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = k0 + PB;                               // first row below the diagonal block = first column of block k+1
    const int row0 = r0 + (int)blockIdx.x * R;            // this CTA's rows
    const int nb1 = min(PB, n - r0);                      // width of block column k+1 = rows of P_top
    const int ntop = (nb1 + R - 1) / R;
    const bool top = (int)blockIdx.x < ntop;
    double* Ak = L + (int64_t)k0 * ld;                    // block column k
    double* An = L + (int64_t)r0 * ld;                    // block column k+1
    int mark = 0;
    auto stamp = [&]() {
        if (PROBE) {
            long long tt;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)::"memory");
            if (tid == 0 && (int)blockIdx.x == probe_cta) dbg[mark] = tt;
            ++mark;
        }
    };
    stamp();

    // ---- X11 (lower triangle, 16-byte pieces) and this CTA's rows of A21 into shared memory ----
    {
        constexpr int PIECES = (PB / 2) * PB / PTHREADS;
        double2 v[PIECES];
#pragma unroll
        for (int u = 0; u < PIECES; ++u) {
            const int idx = tid + u * PTHREADS, p = idx % (PB / 2), c = idx / (PB / 2), r = 2 * p;
            v[u] = make_double2(0.0, 0.0);
            if (r + 1 >= c) v[u] = *reinterpret_cast<const double2*>(Dinv + r + (int64_t)c * ldd);
        }
#pragma unroll
        for (int u = 0; u < PIECES; ++u) {
            const int idx = tid + u * PTHREADS, p = idx % (PB / 2), c = idx / (PB / 2), r = 2 * p;
            if (r + 1 >= c) *reinterpret_cast<double2*>(&sm.X[c * XLD + r]) = v[u];
        }
        constexpr int APT = R * PB / PTHREADS;
        double a[APT];
#pragma unroll
        for (int u = 0; u < APT; ++u) {
            const int idx = tid + u * PTHREADS, rr = idx % R, kk = idx / R;
            a[u] = (row0 + rr < n) ? Ak[row0 + rr + (int64_t)kk * ld] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < APT; ++u) {
            const int idx = tid + u * PTHREADS, rr = idx % R, kk = idx / R;
            sm.A[kk * ALD + rr] = a[u];
        }
    }
    __syncthreads();
    stamp();

    // ---- P = A21 X11^T : P(i,j) = sum_{k <= j} A(i,k) X(j,k).  Warp w owns the column fragments w and 15 - w (together
    // 17 fragment-widths of k: the triangular work is the same for every warp) and all R/8 row fragments ----
    const int q0 = warp, q1 = PB / 8 - 1 - warp;
    double acc[2][RF][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int f = 0; f < RF; ++f) acc[s][f][0] = acc[s][f][1] = 0.0;
    {
        const int kend0 = q0 * 8 + 8, kend1 = q1 * 8 + 8;   // kend0 <= kend1
        const int j0 = q0 * 8 + g, j1 = q1 * 8 + g;
#pragma unroll 4
        for (int kk = 0; kk < kend1; kk += 4) {
            const int k = kk + t;
            double a[RF];
#pragma unroll
            for (int f = 0; f < RF; ++f) a[f] = sm.A[k * ALD + f * 8 + g];
            const double x1 = sm.X[k * XLD + j1];
            const double b1 = (k <= j1) ? x1 : 0.0;
#pragma unroll
            for (int f = 0; f < RF; ++f) dmma_8x8x4(acc[1][f][0], acc[1][f][1], a[f], b1);
            if (kk < kend0) {
                const double x0 = sm.X[k * XLD + j0];
                const double b0 = (k <= j0) ? x0 : 0.0;
#pragma unroll
                for (int f = 0; f < RF; ++f) dmma_8x8x4(acc[0][f][0], acc[0][f][1], a[f], b0);
            }
        }
    }
    stamp();
    __syncthreads();                                       // A21 fully consumed: P replaces it (in shared and global memory)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = (s ? q1 : q0) * 8 + 2 * t;
#pragma unroll
        for (int f = 0; f < RF; ++f) {
            const int rr = f * 8 + g;
            sm.A[c * ALD + rr] = acc[s][f][0];
            sm.A[(c + 1) * ALD + rr] = acc[s][f][1];
            if (row0 + rr < n) {
                Ak[row0 + rr + (int64_t)c * ld] = acc[s][f][0];
                Ak[row0 + rr + (int64_t)(c + 1) * ld] = acc[s][f][1];
            }
        }
    }
    if (top) __threadfence();                              // publishers only: P_top must be visible before the counter moves
    __syncthreads();
    if (top && tid == 0) atomicAdd(counter, 1);
    stamp();

    // ---- old values of this CTA's rows of block column k+1 (independent of P_top: issued before the wait) ----
    // fragment (s, f): rows row0 + 8f + g, columns 8 q_s + 2t, +1 of block column k+1
    double cold[2][RF][2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = (s ? q1 : q0) * 8 + 2 * t;
#pragma unroll
        for (int f = 0; f < RF; ++f) {
            const int row = row0 + f * 8 + g;
            cold[s][f][0] = (row < n && c < nb1) ? __ldcg(An + row + (int64_t)c * ld) : 0.0;
            cold[s][f][1] = (row < n && c + 1 < nb1) ? __ldcg(An + row + (int64_t)(c + 1) * ld) : 0.0;
        }
    }
    if (tid == 0) {
        while (ld_acquire_gpu(counter) < ntop) __nanosleep(100);
    }
    __syncthreads();
    stamp();

    // ---- update: C(i,j) -= sum_k P(i,k) P_top(j,k), K = 128.  B operands (8 rows of P_top per fragment, all 128 k) come
    // straight from L2 into registers: lane (g,t) needs P_top(8q + g, 4 ks + t), ks = 0..31 ----
    // a top CTA only touches the lower triangle of the next diagonal block: fragments entirely above it are skipped
    const int rtop = (int)blockIdx.x * R;                  // this CTA's first row, relative to r0
    bool need[2], load_ok[2];                              // need: warp-uniform (fragment wanted); load_ok: this lane's row of P_top exists
    const double* src[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int q = s ? q1 : q0;
        need[s] = q * 8 < nb1 && !(top && q * 8 > rtop + R - 1);
        load_ok[s] = need[s] && q * 8 + g < nb1;
        src[s] = Ak + r0 + q * 8 + g + (int64_t)t * ld;
    }
    // the 2 x 32 loads of a lane are issued in four groups of 8 k-steps, each group one ahead of the products that
    // consume it: the L2 -> SM transfer (128 KB per CTA) runs under the tensor pipe instead of in front of it
    constexpr int GRP = 8, NGRP = PB / 4 / GRP;
    double bq[2][2][GRP];
    auto load_group = [&](int grp, int buf) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int u = 0; u < GRP; ++u)
                bq[buf][s][u] = load_ok[s] ? __ldcg(src[s] + (int64_t)((grp * GRP + u) * 4) * ld) : 0.0;
    };
    load_group(0, 0);
    stamp();
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int f = 0; f < RF; ++f) acc[s][f][0] = acc[s][f][1] = 0.0;
#pragma unroll
    for (int grp = 0; grp < NGRP; ++grp) {
        if (grp + 1 < NGRP) load_group(grp + 1, (grp + 1) & 1);
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
            const int ks = grp * GRP + u;
            double a[RF];
#pragma unroll
            for (int f = 0; f < RF; ++f) a[f] = sm.A[(ks * 4 + t) * ALD + f * 8 + g];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int f = 0; f < RF; ++f) dmma_8x8x4(acc[s][f][0], acc[s][f][1], a[f], bq[grp & 1][s][u]);
        }
    }
    stamp();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (!need[s]) continue;
        const int c = (s ? q1 : q0) * 8 + 2 * t;
#pragma unroll
        for (int f = 0; f < RF; ++f) {
            const int rr = rtop + f * 8 + g, row = r0 + rr;   // rr: row relative to the next diagonal block
            if (row >= n) continue;
            // strictly above the diagonal of the next diagonal block nothing is written (L's upper triangle stays zero)
            if (c < nb1 && !(top && c > rr)) An[row + (int64_t)c * ld] = cold[s][f][0] - acc[s][f][0];
            if (c + 1 < nb1 && !(top && c + 1 > rr)) An[row + (int64_t)(c + 1) * ld] = cold[s][f][1] - acc[s][f][1];
        }
    }
    stamp();
}

}  // namespace panel
}  // namespace gpirt
