// theta-step contraction on tcgen05 int8 tensor cores with exact integer accumulation (theta_int8.cu).
#pragma once
#include "common.cuh"

namespace gpirt {

struct ThetaInt8 {
    struct Maps;                 // the two TMA tensor maps (opaque here: keeps <cuda.h> out of this header)
    static constexpr int N_CHUNKS = 32;
    int n = 0, m = 0;
    int64_t m_pad = 0, n_pad = 0, c_pad = 0;
    int8_t* yt = nullptr;        // n_pad x m_pad, K-major: responses {+1,-1,0}
    int8_t* yt_abs = nullptr;    // |y| in {0,1} (only when there are missing cells)
    int8_t* Q = nullptr;         // c_pad x m_pad, K-major: digit s of grid row k at row 8 k + s
    double *partial = nullptr, *qscale = nullptr, *oscale = nullptr;
    Maps* maps = nullptr;
    bool ready = false;
    cudaStream_t stream_for_free = nullptr;

    int init(cudaStream_t st, const int8_t* y8, int64_t ldy, int n, int m, bool with_observed_mask);
    // logPt[k + i ldP] (+)= out_factor * sum_j src[k + j ld] * y[i, j]     (k < 1001, i < n; observed_mask: |y|)
    int run(cudaStream_t st, const double* src, int64_t ld, double out_factor, double* logPt, int64_t ldP,
            bool observed_mask, bool accumulate);
    void destroy();
};

// sustained rate of tcgen05.mma.kind::i8 (M128 N256 K32, operands in shared memory) on every SM at once, in 10^12 int8
// operations per second: the measured tensor-pipe ceiling the int8 kernels are judged against
int int8_peak_tops(double* tops, int random_operands = 0);

}  // namespace gpirt
