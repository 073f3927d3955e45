/* TEST INFRASTRUCTURE — not part of the shipped product path.
 *
 * Random-number plumbing for the CPU oracle.  The reference draws everything from R's single global
 * stream (R::rnorm / R::runif, /root/reference/src/mvnormal.h:8, draw-f.cpp:28,35,56, draw-fstar.cpp:27,
 * draw-theta.cpp:27, draw-beta.cpp:22,30, gpirtMCMC.cpp:25).  A GPU sampler cannot reproduce a serial
 * stream, so parity is by *addressed* draws: every draw has an address (sweep, purpose, stream, idx) and a
 * counter-based generator (Philox4x32-10) maps the address to a value.  The CUDA sampler uses the same
 * address scheme (DESIGN.md "RNG addressing"), so oracle and GPU see the same variates.
 *
 * Two sources implement the interface:
 *   - keyed : Philox4x32-10 keyed by a 64-bit seed (optionally recording a tape of everything it hands out,
 *             in consumption order = the reference's serial order);
 *   - tape  : replays a recorded tape sequentially (used to drive the compiled reference sources in
 *             oracle/_ref with exactly the variates the restatement consumed).
 */
#ifndef GPO_RNG_H
#define GPO_RNG_H

#include <cmath>
#include <cstdint>
#include <cstddef>
#include <vector>

namespace gpo {

enum Purpose : uint32_t {
    P_INIT_F_Z  = 0,  /* stream = item j,        idx = respondent i  (normal)  gpirtMCMC.cpp:19-21        */
    P_INIT_BETA = 1,  /* stream = item j,        idx = p in {0,1}    (normal)  gpirtMCMC.cpp:23-27        */
    P_ESS_Z     = 2,  /* stream = item j,        idx = respondent i  (normal)  draw-f.cpp:26              */
    P_ESS_U     = 3,  /* stream = item j,        idx 0: u, 1: eps0, 2+t: t-th shrink redraw (uniform)     */
    P_FSTAR_Z   = 4,  /* stream = item j,        idx = grid point k  (normal)  draw-fstar.cpp:27          */
    P_THETA_U   = 5,  /* stream = respondent i,  idx = 0             (uniform) draw-theta.cpp:27          */
    P_BETA_Z    = 6,  /* stream = item j,        idx = p             (normal)  draw-beta.cpp:22           */
    P_BETA_U    = 7   /* stream = item j,        idx = p             (uniform) draw-beta.cpp:30           */
};

/* Philox4x32-10 (Salmon et al., SC'11).  counter = (c0,c1,c2,c3), key = (k0,k1). */
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

/* 53-bit uniform strictly inside (0,1): (x>>11 + 0.5) * 2^-53 */
inline double u01_from_bits(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* Address -> variates.  counter = (idx', stream, purpose, sweep); normals come in Box-Muller pairs:
 * element idx uses counter idx>>1 and takes the cosine branch when idx is even, the sine branch when odd.
 * Uniforms use counter idx directly and the first 64 output bits. */
inline double keyed_uniform(uint64_t seed, uint32_t sweep, uint32_t purpose, uint32_t stream, uint32_t idx) {
    uint32_t c[4] = {idx, stream, purpose, sweep};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u01_from_bits(c[0], c[1]);
}
inline double keyed_normal(uint64_t seed, uint32_t sweep, uint32_t purpose, uint32_t stream, uint32_t idx) {
    uint32_t c[4] = {idx >> 1, stream, purpose, sweep};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u01_from_bits(c[0], c[1]), u2 = u01_from_bits(c[2], c[3]);
    double r = std::sqrt(-2.0 * std::log(u1));
    const double two_pi = 6.283185307179586476925286766559;
    return (idx & 1u) ? r * std::sin(two_pi * u2) : r * std::cos(two_pi * u2);
}

struct Rng {
    enum Kind { KEYED = 0, TAPE = 1 };
    Kind kind = KEYED;
    uint64_t seed = 0;
    uint32_t sweep = 0;
    bool record = false;
    /* tape storage (recorded or to be replayed): value + kind byte ('n' normal, 'u' uniform) */
    std::vector<double> tape_val;
    std::vector<uint8_t> tape_kind;
    size_t pos = 0;
    int error = 0; /* 1 = tape exhausted, 2 = kind mismatch */
    /* Optional stream maps (keyed mode): a step function run on a SUBSET of items / respondents addresses its variates
     * with the GLOBAL indices of the subset, stream = map[local index] — how the parity tests check sampled items and
     * respondents of a full-size problem against the sampler, whose streams are global indices. */
    std::vector<uint32_t> item_map, resp_map;
    uint32_t map_stream(uint32_t purpose, uint32_t stream) const {
        const std::vector<uint32_t>& m = (purpose == P_THETA_U) ? resp_map : item_map;
        return (!m.empty() && stream < m.size()) ? m[stream] : stream;
    }

    double norm(uint32_t purpose, uint32_t stream, uint32_t idx) {
        if (kind == TAPE) return pop('n');
        stream = map_stream(purpose, stream);
        double v = keyed_normal(seed, sweep, purpose, stream, idx);
        if (record) { tape_val.push_back(v); tape_kind.push_back('n'); }
        return v;
    }
    double unif(uint32_t purpose, uint32_t stream, uint32_t idx) {
        if (kind == TAPE) return pop('u');
        stream = map_stream(purpose, stream);
        double v = keyed_uniform(seed, sweep, purpose, stream, idx);
        if (record) { tape_val.push_back(v); tape_kind.push_back('u'); }
        return v;
    }
    double pop(uint8_t want) {
        if (pos >= tape_val.size()) { error = 1; return std::nan(""); }
        if (tape_kind[pos] != want) { error = 2; }
        return tape_val[pos++];
    }
};

/* R nmath restatements (R >= 3.4, src/nmath/{rnorm,runif,dnorm,plogis}.c semantics). */
inline double r_rnorm(double mu, double sigma, double z) { return mu + sigma * z; }
inline double r_runif(double a, double b, double u) { return a + (b - a) * u; }
inline double r_dnorm_log(double x, double mu, double sigma) {
    const double LN_SQRT_2PI = 0.918938533204672741780329736406;
    double t = (x - mu) / sigma;
    t = std::fabs(t);
    return -(LN_SQRT_2PI + 0.5 * t * t + std::log(sigma));
}
inline double r_plogis(double x) { return 1.0 / (1.0 + std::exp(-x)); }

} // namespace gpo
#endif
