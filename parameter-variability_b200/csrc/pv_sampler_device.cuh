// Device-resident sampler loop: adaptive Metropolis inside adaptive parallel tempering for Q problems x C temperatures
// (reference: pyPESTO's AdaptiveParallelTemperingSampler(AdaptiveMetropolisSampler, 4), 5000 samples per problem, set up
// at src/parvar/optimization/petab_optimization.py:158-173).  One iteration is
//     pv_sampler_propose_kernel  ->  the batched objective (pv_logp: C*Q*G trajectories)  ->  pv_sampler_step_kernel
// queued on one stream and captured ONCE into a CUDA graph that is replayed per iteration: chain state, proposal
// covariances, temperature ladder and the trace live in HBM, nothing of an iteration touches the host (the host-loop
// version paid H2D + D2H + ~10 launches' latency per iteration, a 50x tax on the small latency-bound launches).  Random
// numbers are the caller's, uploaded in blocks of iterations in the order optimization/sampler.py draws them, so the
// device loop walks the same chains as the numpy statement of the algorithm (up to round-off).
// The arithmetic of an iteration is the same code as the host bookkeeping (pv_sampler_host.cuh), compiled for the device.
#pragma once
#include <cuda_runtime.h>
#include "pv_sampler_host.cuh"

struct pv_sampler_dev {
    int64_t Q;                     // problems
    int C, G, NP, dim, S;          // temperatures, groups, parameters per group, NP * G, max(C - 1, 1)
    int64_t n_trace, rng_block;    // trace rows (n_samples + 1), iterations per uploaded block of random numbers
    double *x, *llh, *lpr, *betas, *mean_hist, *cov_hist, *cov_scale, *cov, *n_acc, *n_swap, *n_swap_try;   // [Q][C]...
    double *x_new, *lpr_new;       // [Q][C][dim], [Q][C]
    unsigned char* inside;         // [Q][C]
    double* P;                     // [NP][C*Q*G]  objective input: trajectory (c*Q + q)*G + g <- x_eval[q][c][p*G + g]
    double* seg;                   // [C*Q]        objective output: log-likelihood of x_eval[q][c] at index c*Q + q
    const double *lb, *ub;         // [dim]
    const double *prior_loc, *prior_scale;   // [Q][dim], NaN = no prior on that parameter
    const double *xi, *u_acc, *u_swap;       // [rng_block][Q][C][dim], [rng_block][Q][C], [rng_block][Q][S]
    double *trace_x, *trace_lp, *trace_pr;   // [n_trace][Q][C][dim], [n_trace][Q][C] x 2
    long long* it;                 // iteration being executed (1-based; 0 = the initial evaluation)
    unsigned int* done;            // blocks of the step kernel that have finished (last one advances `it`)
    unsigned long long* n_bad_cov; // proposals skipped because a covariance was not positive definite
    pv_sampler_options o;
};

// one thread per (problem, temperature): proposal, bounds, objective input, log-prior of the evaluated point
__global__ void pv_sampler_propose_kernel(const pv_sampler_dev s) {
    constexpr double LOG_2PI = 1.8378770664093454835606594728112;
    const int64_t qc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qc >= s.Q * s.C) return;
    const int64_t q = qc / s.C;
    const int c = (int)(qc % s.C), dim = s.dim;
    const long long it = *s.it;
    const double* x = s.x + qc * dim;
    double* xn = s.x_new + qc * dim;
    bool in = true;
    if (it == 0) {
        for (int i = 0; i < dim; ++i) xn[i] = x[i];
    } else {
        const int64_t slot = (it - 1) % s.rng_block;
        const double* xi = s.xi + (slot * s.Q * s.C + qc) * dim;
        double low[pv_sampler_math::MAXD * pv_sampler_math::MAXD];
        const bool ok = pv_sampler_math::cholesky(dim, s.cov + qc * dim * dim, low);
        if (!ok) atomicAdd(s.n_bad_cov, 1ull);
        for (int i = 0; i < dim; ++i) {
            double acc = 0.0;
            if (ok)
                for (int j = 0; j <= i; ++j) acc += low[i * dim + j] * xi[j];
            const double v = x[i] + acc;
            xn[i] = v;
            in = in && v >= s.lb[i] && v <= s.ub[i];
        }
    }
    s.inside[qc] = in;
    // the objective is asked for x_eval = inside ? x_new : x; its log-prior (PEtab parameterScaleNormal on the lin scale)
    const int64_t BT = (int64_t)s.C * s.Q * s.G;
    double z2 = 0.0, cst = 0.0;
    for (int i = 0; i < dim; ++i) {
        const double v = in ? xn[i] : x[i];
        const int p = i / s.G, g = i % s.G;
        s.P[(int64_t)p * BT + ((int64_t)c * s.Q + q) * s.G + g] = v;
        const double loc = s.prior_loc[q * dim + i], sc = s.prior_scale[q * dim + i];
        if (loc == loc) {
            const double z = (v - loc) / sc;
            z2 += z * z;
            cst += LOG_2PI + 2.0 * log(sc);
        }
    }
    s.lpr_new[qc] = -0.5 * z2 - 0.5 * cst;
}

// blockDim.x = C * (problems per block): all temperatures of a problem sit in one block
__global__ void pv_sampler_step_kernel(const pv_sampler_dev s) {
    const int C = s.C, dim = s.dim, S = s.S;
    const int qpb = blockDim.x / C;
    const int64_t q = (int64_t)blockIdx.x * qpb + threadIdx.x / C;
    const int c = threadIdx.x % C;
    const bool live = threadIdx.x < qpb * C && q < s.Q;
    const long long it = *s.it;
    const int64_t qc = q * C + c;
    const double ninf = -INFINITY;
    bool* swapped = reinterpret_cast<bool*>(0);
    extern __shared__ unsigned char sh_swapped[];     // [qpb][S]
    (void)swapped;
    if (live) {
        double lnew = s.seg[(int64_t)c * s.Q + q];
        if (!(fabs(lnew) <= 1.7e308)) lnew = ninf;    // failed integration (NaN) or overflow: a failed evaluation
        if (it == 0) {
            s.llh[qc] = lnew;
            s.lpr[qc] = s.lpr_new[qc];
        } else {
            const int64_t slot = (it - 1) % s.rng_block;
            if (!s.inside[qc]) lnew = ninf;
            double log_acc = s.betas[qc] * (lnew - s.llh[qc]) + (s.lpr_new[qc] - s.lpr[qc]);
            log_acc = (log_acc != log_acc) ? ninf : fmin(log_acc, 0.0);
            const bool accept = log(s.u_acc[slot * s.Q * C + qc]) < log_acc;
            double* x = s.x + qc * dim;
            if (accept) {
                const double* xn = s.x_new + qc * dim;
                for (int i = 0; i < dim; ++i) x[i] = xn[i];
                s.llh[qc] = lnew;
                s.lpr[qc] = s.lpr_new[qc];
            }
            s.n_acc[qc] += accept;
            if (it > s.o.threshold_sample) {   // Haario et al.
                const double rate = pow((double)it, -s.o.decay_constant);
                const double scale_gain = 1.0 / pow((double)(it + 1), s.o.decay_constant);
                double dx[pv_sampler_math::MAXD];
                double* mh = s.mean_hist + qc * dim;
                for (int i = 0; i < dim; ++i) {
                    mh[i] = (1 - rate) * mh[i] + rate * x[i];
                    dx[i] = x[i] - mh[i];
                }
                double* ch = s.cov_hist + qc * dim * dim;
                for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) ch[i * dim + j] = (1 - rate) * ch[i * dim + j] + rate * (dx[i] * dx[j]);
                s.cov_scale[qc] *= exp((exp(log_acc) - s.o.target_acceptance_rate) * scale_gain);
                double cv[pv_sampler_math::MAXD * pv_sampler_math::MAXD];
                for (int i = 0; i < dim * dim; ++i) cv[i] = s.cov_scale[qc] * ch[i];
                pv_sampler_math::regularize(dim, cv, s.o.reg_factor);
                double* out = s.cov + qc * dim * dim;
                for (int i = 0; i < dim * dim; ++i) out[i] = cv[i];
            }
        }
    }
    __syncthreads();
    if (live && c == 0 && it > 0) {
        const int64_t slot = (it - 1) % s.rng_block;
        unsigned char* sw = sh_swapped + (threadIdx.x / C) * S;
        // ---- swaps, hottest pair first ----
        for (int k = C - 1; k >= 1; --k) {
            const int64_t a = q * C + (k - 1), b = q * C + k;
            double p_swap = (s.betas[a] - s.betas[b]) * (s.llh[b] - s.llh[a]);
            if (p_swap != p_swap) p_swap = ninf;
            const bool doit = log(s.u_swap[(slot * s.Q + q) * S + (k - 1)]) < p_swap;
            s.n_swap_try[q * S + (k - 1)] += 1;
            s.n_swap[q * S + (k - 1)] += doit;
            sw[k - 1] = doit;
            if (doit) {
                for (int i = 0; i < dim; ++i) {
                    const double tmp = s.x[a * dim + i];
                    s.x[a * dim + i] = s.x[b * dim + i];
                    s.x[b * dim + i] = tmp;
                }
                double tmp = s.llh[a];
                s.llh[a] = s.llh[b];
                s.llh[b] = tmp;
                tmp = s.lpr[a];
                s.lpr[a] = s.lpr[b];
                s.lpr[b] = tmp;
            }
        }
        // ---- ladder adaptation (Vousden et al. 2016) ----
        if (s.o.adapt_betas && C > 2) {
            const double kappa = s.o.nu / ((double)(it + 1) + s.o.nu) / s.o.eta;
            const double t0 = 1.0 / s.betas[q * C];
            double run = t0, prev = t0;
            for (int j = 0; j < C - 2; ++j) {
                const double tj1 = 1.0 / s.betas[q * C + j + 1];
                const double ds = kappa * ((double)sw[j] - (double)sw[j + 1]);
                const double dt = (tj1 - prev) * exp(ds);
                prev = tj1;
                run += dt;
                s.betas[q * C + j + 1] = 1.0 / run;
            }
        }
    }
    __syncthreads();
    if (live) {
        const int64_t row = (int64_t)it * s.Q * C + qc;
        for (int i = 0; i < dim; ++i) s.trace_x[row * dim + i] = s.x[qc * dim + i];
        s.trace_lp[row] = -(s.llh[qc] + s.lpr[qc]);
        s.trace_pr[row] = -s.lpr[qc];
    }
    // the last block to finish advances the iteration counter (every block has read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(s.done, 1u) == gridDim.x - 1) {
            *s.done = 0;
            *s.it = it + 1;
        }
    }
}
