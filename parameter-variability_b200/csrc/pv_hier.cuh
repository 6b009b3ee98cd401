// Device-resident epilogue of the hierarchical (population) objective of BASELINE configs 3-4:
// the per-individual log-likelihoods and gradients the sensitivity kernel leaves in HBM are folded with the log-normal
// population density into ONE small buffer per evaluation,
//     red[c] = ( sum_i [loglik_ci + log LogNormal(theta_ci; mu_c, omega_c)],  d/d mu_c[P],  d/d omega_c[P] ),
// which is the send buffer of the single all-reduce (NCCL, same stream) when the individuals of a chain are sharded over
// GPUs; per-individual gradient blocks never leave their GPU.  Nothing of an evaluation touches the host.  The
// reference has no analogue: its only parallel strategy is a process pool over problems
// (src/parvar/optimization/petab_optimization.py:381-391).
#pragma once
#include <cuda_runtime.h>
#include "pv_common.cuh"

#define PV_HIER_MAX_P 8          // = PV_MAX_COLS
#define PV_HIER_THREADS 256

struct pv_hier_params {
    int n_chains;                // C
    int64_t n_local;             // L individuals of every chain on this GPU
    int n_par;                   // P per-individual parameters (columns)
    const double* theta;         // [P][C*L]   (chain-major columns: c*L + i)
    const double* mu;            // [C][P]     log-scale population means
    const double* omega;         // [C][P]     log-scale population sds
    const double* logp;          // [C*L]      per-individual log-likelihood (NaN = failed integration)
    const double* grad;          // [NS][C*L]  d loglik / d theta in the model's compiled sensitivity order
    double* dtheta;              // [P][C*L]   out: d log posterior / d theta_ci  (likelihood + population density)
    double* red;                 // [C][1+2P]  out: this rank's partial sums
    int sens_row[PV_HIER_MAX_P]; // column p -> row of grad
};

// One CTA per chain; every thread walks the chain's individuals with a fixed stride and the block reduces in a fixed
// tree, so the sums are bit-reproducible from launch to launch (and independent of how many GPUs share the box).
__global__ void __launch_bounds__(PV_HIER_THREADS) pv_hier_epilogue_kernel(const pv_hier_params p) {
    constexpr double LOG_2PI = 1.8378770664093454835606594728112;
    const int c = blockIdx.x;
    const int P = p.n_par;
    const int64_t L = p.n_local, B = (int64_t)p.n_chains * L;
    double acc[1 + 3 * PV_HIER_MAX_P];     // logp | sum log theta_p | sum z_p | sum z_p^2
#pragma unroll
    for (int k = 0; k < 1 + 3 * PV_HIER_MAX_P; ++k) acc[k] = 0.0;
    double mu[PV_HIER_MAX_P], iw[PV_HIER_MAX_P];
#pragma unroll
    for (int q = 0; q < PV_HIER_MAX_P; ++q) {
        mu[q] = q < P ? p.mu[c * P + q] : 0.0;
        iw[q] = q < P ? 1.0 / p.omega[c * P + q] : 0.0;
    }
    for (int64_t i = threadIdx.x; i < L; i += PV_HIER_THREADS) {
        const int64_t b = (int64_t)c * L + i;
        acc[0] += p.logp[b];
#pragma unroll
        for (int q = 0; q < PV_HIER_MAX_P; ++q)
            if (q < P) {
                const double th = p.theta[(int64_t)q * B + b];
                const double lt = log(th);                      // theta <= 0: NaN, poisons the chain's logp like a failed solve
                const double z = (lt - mu[q]) * iw[q];
                acc[1 + q] += lt;
                acc[1 + PV_HIER_MAX_P + q] += z;
                acc[1 + 2 * PV_HIER_MAX_P + q] = fma(z, z, acc[1 + 2 * PV_HIER_MAX_P + q]);
                // d/d theta of log LogNormal(theta; mu, omega) = -(1 + z/omega)/theta
                p.dtheta[(int64_t)q * B + b] = p.grad[(int64_t)p.sens_row[q] * B + b] - fma(z, iw[q], 1.0) / th;
            }
    }
    __shared__ double sh[PV_HIER_THREADS / 32][1 + 3 * PV_HIER_MAX_P];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 1 + 3 * PV_HIER_MAX_P; ++k) {
        double v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 1 + 3 * PV_HIER_MAX_P) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < PV_HIER_THREADS / 32; ++w) v += sh[w][threadIdx.x];
        sh[0][threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* out = p.red + (int64_t)c * (1 + 2 * P);
        double pop = -0.5 * LOG_2PI * (double)L * P;
        for (int q = 0; q < P; ++q) {
            const double s_lt = sh[0][1 + q], s_z = sh[0][1 + PV_HIER_MAX_P + q], s_zz = sh[0][1 + 2 * PV_HIER_MAX_P + q];
            pop += -s_lt + (double)L * log(iw[q]) - 0.5 * s_zz;
            out[1 + q] = s_z * iw[q];
            out[1 + P + q] = (s_zz - (double)L) * iw[q];
        }
        out[0] = sh[0][0] + pop;
    }
}

struct pv_hier_prior {
    double m[PV_HIER_MAX_P], s[PV_HIER_MAX_P], h[PV_HIER_MAX_P];   // mu_p ~ Normal(m_p, s_p), omega_p ~ HalfNormal(h_p)
};

// Hyper-priors, added once AFTER the cross-rank reduction, in place: red[c] becomes (log posterior, d/d mu, d/d omega).
__global__ void pv_hier_finish_kernel(int C, int P, double* red, const double* mu, const double* omega, const pv_hier_prior pr) {
    constexpr double LOG_2PI = 1.8378770664093454835606594728112;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double* r = red + (int64_t)c * (1 + 2 * P);
    double lp = r[0];
    for (int q = 0; q < P; ++q) {
        const double zm = (mu[c * P + q] - pr.m[q]) / pr.s[q];
        const double om = omega[c * P + q], h = pr.h[q];
        lp += -0.5 * (LOG_2PI + 2.0 * log(pr.s[q])) - 0.5 * zm * zm;
        lp += 0.5 * log(2.0 / 3.14159265358979323846) - log(h) - 0.5 * (om / h) * (om / h);
        r[1 + q] -= zm / pr.s[q];
        r[1 + P + q] -= om / (h * h);
    }
    r[0] = lp;
}
