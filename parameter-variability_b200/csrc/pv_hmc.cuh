// Device-resident Hamiltonian Monte Carlo on the hierarchical (population) posterior of BASELINE config 3 - "the
// log-likelihood and its gradient that the MCMC / NUTS sampler calls at every step" (north_star) with the sampler's own
// state on the GPU as well: position, momentum, gradient, step size, mass matrix and the kept samples live in HBM, one
// leapfrog step is
//     pv_hmc_drift_kernel -> pv_hier_logp_grad (sensitivity kernel + epilogue) [-> all-reduce] -> pv_hier_finish -> pv_hmc_kick_kernel
// queued on one stream without a host synchronisation, and the accept / step-size / mass adaptation of an iteration is one
// more kernel.  Same algorithm, same random numbers and the same arithmetic order per chain as the numpy statement
// optimization/hmc.py::hmc (static trajectory length with jitter, dual averaging, one-window diagonal mass adaptation), which
// paid H2D of theta and D2H of all per-individual gradients at every leapfrog step.  All positive quantities are sampled on
// the log scale: x = (log theta_ci, mu_c, log omega_c), Jacobian terms included here.
// The reference has no gradient-based sampler of its own (pyPESTO's adaptive Metropolis, petab_optimization.py:158-173); this
// is the north-star's consumer of the logp + gradient op.
// One CTA per chain; every reduction is a fixed-order block reduction, so a run reproduces itself bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/parvar_b200.h"   // pv_hmc_state

#define PV_HMC_THREADS 256

__device__ __forceinline__ double pv_hmc_block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();                 // sh may still be read by the previous call
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < PV_HMC_THREADS / 32; ++k) tot += sh[k];      // same order in every thread
    return tot;
}

// momenta from the iteration's normals, kinetic energy at the start, trial point <- state
__global__ void pv_hmc_begin_kernel(const pv_hmc_state s, const double* __restrict__ z) {
    __shared__ double sh[PV_HMC_THREADS / 32];
    const int c = blockIdx.x, P = s.n_par;
    const int64_t L = s.n_local, CL = (int64_t)s.n_chains * L, dim = s.n_total * P + 2 * P;
    const double* zc = z + (int64_t)c * dim;
    double K = 0.0;
    for (int64_t e = threadIdx.x; e < (int64_t)P * L; e += blockDim.x) {
        const int p = (int)(e / L);
        const int64_t i = e - (int64_t)p * L, k = (int64_t)p * CL + (int64_t)c * L + i;
        const double im = s.im_lth[k], mom = zc[(s.first + i) * P + p] / sqrt(im);
        s.p_lth[k] = mom;
        K += 0.5 * mom * mom * im;
        s.lth_q[k] = s.lth[k];
        s.g_lth_q[k] = s.g_lth[k];
    }
    if (threadIdx.x < 2 * P) {
        const int j = threadIdx.x, k = c * 2 * P + j;
        const double im = s.im_hyp[k], mom = zc[s.n_total * P + j] / sqrt(im);
        s.p_hyp[k] = mom;
        if (s.own_hyper) K += 0.5 * mom * mom * im;
        s.g_hyp_q[k] = s.g_hyp[k];
        if (j < P)
            s.mu_q[c * P + j] = s.mu[c * P + j];
        else
            s.lom_q[c * P + j - P] = s.lom[c * P + j - P];
    }
    K = pv_hmc_block_sum(K, sh);
    if (threadIdx.x == 0) {
        s.energy[c * 4 + 0] = K;
        s.energy[c * 4 + 1] = 0.0;
        s.energy[c * 4 + 2] = 0.0;
        s.energy[c * 4 + 3] = 0.0;
        s.ok[c] = 1;
    }
}

// half kick with the gradient at the trial point, full drift, objective inputs theta = exp(log theta), omega = exp(log omega).
// scale = 0: no move (evaluation of the start point)
__global__ void pv_hmc_drift_kernel(const pv_hmc_state s, double scale) {
    const int c = blockIdx.x, P = s.n_par;
    const int64_t L = s.n_local, CL = (int64_t)s.n_chains * L;
    const double eps = scale * s.eps[c];
    for (int64_t e = threadIdx.x; e < (int64_t)P * L; e += blockDim.x) {
        const int p = (int)(e / L);
        const int64_t k = (int64_t)p * CL + (int64_t)c * L + (e - (int64_t)p * L);
        const double pq = s.p_lth[k] + 0.5 * eps * s.g_lth_q[k];
        const double xq = s.lth_q[k] + eps * s.im_lth[k] * pq;
        s.p_lth[k] = pq;
        s.lth_q[k] = xq;
        s.theta[k] = exp(xq);
    }
    if (threadIdx.x < 2 * P) {
        const int j = threadIdx.x, k = c * 2 * P + j;
        const double pq = s.p_hyp[k] + 0.5 * eps * s.g_hyp_q[k];
        s.p_hyp[k] = pq;
        if (j < P) {
            s.mu_q[c * P + j] += eps * s.im_hyp[k] * pq;
        } else {
            const double xq = s.lom_q[c * P + j - P] + eps * s.im_hyp[k] * pq;
            s.lom_q[c * P + j - P] = xq;
            s.omega[c * P + j - P] = exp(xq);
        }
    }
}

// gradient of the transformed density at the new trial point (chain rule of theta = exp(.), omega = exp(.), + 1 from the
// Jacobian), second half kick; a non-finite evaluation marks the trajectory and kicks with zero.  last != 0: kinetic energy at
// the end and the Jacobian term sum(log theta) + sum(log omega) of the trial point (local partials; the caller all-reduces them)
__global__ void pv_hmc_kick_kernel(const pv_hmc_state s, double scale, int last) {
    __shared__ double sh[PV_HMC_THREADS / 32];
    const int c = blockIdx.x, P = s.n_par;
    const int64_t L = s.n_local, CL = (int64_t)s.n_chains * L;
    const double eps = scale * s.eps[c];
    const double* red = s.red + (int64_t)c * (1 + 2 * P);
    bool bad = !isfinite(red[0]);
    for (int j = 0; j < 2 * P; ++j) bad = bad || !isfinite(red[1 + j]);       // (all-reduced: every rank decides alike)
    if (bad && threadIdx.x == 0) s.ok[c] = 0;
    double K = 0.0, J = 0.0;
    for (int64_t e = threadIdx.x; e < (int64_t)P * L; e += blockDim.x) {
        const int p = (int)(e / L);
        const int64_t k = (int64_t)p * CL + (int64_t)c * L + (e - (int64_t)p * L);
        const double g = bad ? 0.0 : fma(s.dtheta[k], s.theta[k], 1.0);
        const double pq = s.p_lth[k] + 0.5 * eps * g;
        s.g_lth_q[k] = g;
        s.p_lth[k] = pq;
        K += 0.5 * pq * pq * s.im_lth[k];
        J += s.lth_q[k];
    }
    if (threadIdx.x < 2 * P) {
        const int j = threadIdx.x, k = c * 2 * P + j;
        const double g = bad ? 0.0 : (j < P ? red[1 + j] : fma(red[1 + j], s.omega[c * P + j - P], 1.0));
        const double pq = s.p_hyp[k] + 0.5 * eps * g;
        s.g_hyp_q[k] = g;
        s.p_hyp[k] = pq;
        if (s.own_hyper) {
            K += 0.5 * pq * pq * s.im_hyp[k];
            if (j >= P) J += s.lom_q[c * P + j - P];
        }
    }
    if (last) {
        K = pv_hmc_block_sum(K, sh);
        J = pv_hmc_block_sum(J, sh);
        if (threadIdx.x == 0) {
            s.energy[c * 4 + 1] = K;
            s.energy[c * 4 + 2] = J;
        }
    }
}

// Metropolis test on the (all-reduced) energies, state <- trial point where accepted, dual averaging of the step size and
// Welford moments for the mass matrix during warm-up, kept sample otherwise.  mode 0: adopt the trial point unconditionally
// (start of a run); mode 1: warm-up iteration `it`; mode 2: sampling iteration, sample slot `it - n_warmup`.
__global__ void pv_hmc_accept_kernel(const pv_hmc_state s, int mode, int64_t it, int64_t n_warmup, const double* __restrict__ u,
                                     double target_accept, double max_energy_error, int adapt_mass, int64_t w_n) {
    const int c = blockIdx.x, P = s.n_par;
    const int64_t L = s.n_local, CL = (int64_t)s.n_chains * L;
    const double lpq = s.red[(int64_t)c * (1 + 2 * P)] + s.energy[c * 4 + 2];
    const bool ok = s.ok[c] != 0;
    double a = 1.0;
    bool accept = true, diverged = false;
    if (mode != 0) {
        const double h0 = -s.lp[c] + s.energy[c * 4 + 0], h1 = -lpq + s.energy[c * 4 + 1];
        const double dh = ok ? h0 - h1 : -INFINITY;
        diverged = !ok || dh < -max_energy_error || dh != dh;
        a = diverged ? 0.0 : fmin(1.0, exp(fmin(dh, 0.0)));
        accept = u[c] < a;
    }
    __syncthreads();                 // every thread has read lp[c] before thread 0 replaces it
    const bool welford = mode == 1 && adapt_mass && w_n > 0;         // w_n: this iteration's 1-based count of Welford updates
    const bool finish_mass = mode == 1 && it == n_warmup - 1 && adapt_mass && w_n > 10;
    for (int64_t e = threadIdx.x; e < (int64_t)P * L; e += blockDim.x) {
        const int p = (int)(e / L);
        const int64_t k = (int64_t)p * CL + (int64_t)c * L + (e - (int64_t)p * L);
        if (accept) {
            s.lth[k] = s.lth_q[k];
            s.g_lth[k] = s.g_lth_q[k];
        }
        const double x = s.lth[k];
        if (welford) {
            const double d = x - s.w_mean_lth[k], m = s.w_mean_lth[k] + d / (double)w_n;
            s.w_mean_lth[k] = m;
            s.w_m2_lth[k] += d * (x - m);
        }
        if (finish_mass) {
            const double var = s.w_m2_lth[k] / (double)(w_n - 1);
            s.im_lth[k] = ((double)w_n / ((double)w_n + 5.0)) * var + 1e-3 * (5.0 / ((double)w_n + 5.0));      // Stan's shrinkage
        }
        if (mode == 2 && s.samples_lth) s.samples_lth[(it - n_warmup) * (int64_t)P * CL + k] = x;
    }
    if (threadIdx.x < 2 * P) {
        const int j = threadIdx.x, k = c * 2 * P + j;
        double* xs = j < P ? s.mu + c * P + j : s.lom + c * P + j - P;
        const double* xq = j < P ? s.mu_q + c * P + j : s.lom_q + c * P + j - P;
        if (accept) {
            *xs = *xq;
            s.g_hyp[k] = s.g_hyp_q[k];
        }
        const double x = *xs;
        if (welford) {
            const double d = x - s.w_mean_hyp[k], m = s.w_mean_hyp[k] + d / (double)w_n;
            s.w_mean_hyp[k] = m;
            s.w_m2_hyp[k] += d * (x - m);
        }
        if (finish_mass) {
            const double var = s.w_m2_hyp[k] / (double)(w_n - 1);
            s.im_hyp[k] = ((double)w_n / ((double)w_n + 5.0)) * var + 1e-3 * (5.0 / ((double)w_n + 5.0));
        }
        if (mode == 2 && s.samples_hyp) s.samples_hyp[((it - n_warmup) * s.n_chains + c) * 2 * P + j] = x;
    }
    if (threadIdx.x == 0) {
        if (accept) s.lp[c] = lpq;
        if (mode == 1) {
            // dual averaging (Hoffman & Gelman 2014, Algorithm 5): da = (mu, log_eps_bar, h_bar)
            constexpr double gamma = 0.05, t0 = 10.0, kappa = 0.75;
            double* da = s.da + c * 4;
            const double m = (double)(it + 1);
            da[2] = (1.0 - 1.0 / (m + t0)) * da[2] + (target_accept - a) / (m + t0);
            const double log_eps = da[0] - sqrt(m) / gamma * da[2];
            const double eta = pow(m, -kappa);
            da[1] = eta * log_eps + (1.0 - eta) * da[1];
            s.eps[c] = (it == n_warmup - 1) ? exp(da[1]) : exp(log_eps);
        } else if (mode == 2) {
            s.samples_lp[(it - n_warmup) * s.n_chains + c] = s.lp[c];
            s.acc_sum[c] += a;
            s.n_div[c] += diverged ? 1 : 0;
        }
    }
}
