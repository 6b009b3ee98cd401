// numpy's frozen legacy random stream (np.random.RandomState: MT19937 + polar Box-Muller with a cached second deviate),
// restated on the host so that the sampler can draw its blocks of random numbers in compiled code, straight into the
// pinned staging buffers, with the interpreter lock released - and still walk exactly the chain the numpy statement of
// the algorithm (optimization/sampler.py::sample) walks from the same seed.  The reference's own path uses this very
// stream for sampling and noise (np.random.seed / np.random.lognormal / np.random.normal,
// src/parvar/experiments/petab_factory.py:103-118, 201-216).
// Published algorithms: Matsumoto & Nishimura (1998) MT19937; Marsaglia's polar method.  State import / export through
// RandomState.get_state() / set_state(), so seeding stays numpy's.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/parvar_b200.h"   // pv_mt19937_state

namespace pv_rng {
inline void regenerate(pv_mt19937_state* s) {
    constexpr int N = 624, M = 397;
    constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    uint32_t* mt = s->key;
    int kk = 0;
    uint32_t y;
    for (; kk < N - M; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < N - 1; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
    mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    s->pos = 0;
}
inline uint32_t next32(pv_mt19937_state* s) {
    if (s->pos >= 624) regenerate(s);
    uint32_t y = s->key[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
// 53-bit double in [0, 1): numpy's random_double / RandomState.uniform()
inline double next_double(pv_mt19937_state* s) {
    const int32_t a = (int32_t)(next32(s) >> 5), b = (int32_t)(next32(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
// RandomState.normal(): polar method, the second deviate of a pair is kept for the next call
inline double next_gauss(pv_mt19937_state* s) {
    if (s->has_gauss) {
        const double t = s->gauss;
        s->has_gauss = 0;
        s->gauss = 0.0;
        return t;
    }
    double x1, x2, r2;
    do {
        x1 = 2.0 * next_double(s) - 1.0;
        x2 = 2.0 * next_double(s) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = sqrt(-2.0 * log(r2) / r2);
    s->gauss = f * x1;
    s->has_gauss = 1;
    return f * x2;
}
// The draws of n_iter sampler iterations in the numpy loop's order: per iteration normal(size=(Q, C, dim)),
// uniform(size=(Q, C)), then for k = C-1 .. 1 uniform(size=Q) -> u_swap[:, k-1].
inline void fill_sampler_block(pv_mt19937_state* s, int64_t n_iter, int64_t Q, int C, int dim, double* xi, double* u_acc,
                               double* u_swap) {
    const int S = C > 1 ? C - 1 : 1;
    for (int64_t it = 0; it < n_iter; ++it) {
        double* x = xi + it * Q * C * dim;
        for (int64_t i = 0; i < Q * C * dim; ++i) x[i] = next_gauss(s);
        double* ua = u_acc + it * Q * C;
        for (int64_t i = 0; i < Q * C; ++i) ua[i] = next_double(s);
        double* us = u_swap + it * Q * S;
        if (C == 1)
            for (int64_t q = 0; q < Q; ++q) us[q] = 0.0;
        for (int k = C - 1; k >= 1; --k)
            for (int64_t q = 0; q < Q; ++q) us[q * S + (k - 1)] = next_double(s);
    }
}
}  // namespace pv_rng
