// Host-side bookkeeping of the batched sampler: adaptive Metropolis inside adaptive parallel tempering for Q problems x
// C temperatures at once (reference: pyPESTO's AdaptiveParallelTemperingSampler(AdaptiveMetropolisSampler, 4) as set
// up at src/parvar/optimization/petab_optimization.py:158-173, 5000 samples per problem).
//
// The likelihood of every proposal is the GPU objective (one launch per iteration for all problems and temperatures);
// what is here is the part around it - proposal from the adapted covariance, accept / reject, Haario adaptation,
// temperature swaps, ladder adaptation, trace - which in numpy cost 1.5 us per chain and iteration (65 s per
// 2160-problem model of the reference's study, more than the GPU time of the two small models).  Plain C++ over
// caller-owned host arrays; random numbers come from the caller (numpy's RandomState, same stream order as
// optimization/sampler.py::sample, so both implementations walk the same chain up to round-off).
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/parvar_b200.h"   // pv_sampler_options

#if defined(__CUDACC__)
#define PV_SAMPLER_HD __host__ __device__ inline
#else
#define PV_SAMPLER_HD inline
#endif

namespace pv_sampler_math {
constexpr int MAXD = 16;

// in-place Cholesky of the symmetric a [n][n] (row-major, lower triangle used); false if a pivot is not > 0
PV_SAMPLER_HD bool cholesky(int n, const double* a, double* low) {
    for (int i = 0; i < n * n; ++i) low[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double d = a[j * n + j];
        for (int k = 0; k < j; ++k) d -= low[j * n + k] * low[j * n + k];
        if (!(d > 0.0)) return false;
        const double dj = sqrt(d);
        low[j * n + j] = dj;
        for (int i = j + 1; i < n; ++i) {
            double s = a[i * n + j];
            for (int k = 0; k < j; ++k) s -= low[i * n + k] * low[j * n + k];
            low[i * n + j] = s / dj;
        }
    }
    return true;
}

// cyclic Jacobi eigen-decomposition of the symmetric a [n][n]: a = V diag(w) V^T (V row-major, columns = vectors)
PV_SAMPLER_HD void jacobi_eigh(int n, double* a, double* w, double* V) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) (i == j ? diag : off) += a[i * n + j] * a[i * n + j];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {   // columns p, q of a
                    const double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = c * akp - s * akq;
                    a[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {   // rows p, q of a
                    const double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = c * apk - s * aqk;
                    a[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = a[i * n + i];
}

// optimization/sampler.py::_regularize for one matrix: symmetrise; eigenvalues >= max(reg, 1e-12 lambda_max).
// The cheap test first: cov - f I positive definite for f = max(reg, 1e-12 trace) >= that floor.
PV_SAMPLER_HD void regularize(int n, double* cov, double reg) {
    double tr = 0.0;
    for (int i = 0; i < n; ++i) {
        tr += cov[i * n + i];
        for (int j = i + 1; j < n; ++j) cov[i * n + j] = cov[j * n + i] = 0.5 * (cov[i * n + j] + cov[j * n + i]);
    }
    const double bound = fmax(reg, 1e-12 * fabs(tr));
    double shifted[MAXD * MAXD], low[MAXD * MAXD];
    for (int i = 0; i < n * n; ++i) shifted[i] = cov[i];
    for (int i = 0; i < n; ++i) shifted[i * n + i] -= bound;
    if (cholesky(n, shifted, low)) return;
    double a[MAXD * MAXD], w[MAXD], V[MAXD * MAXD];
    for (int i = 0; i < n * n; ++i) a[i] = cov[i];
    jacobi_eigh(n, a, w, V);
    double wmax = w[0];
    for (int i = 1; i < n; ++i) wmax = fmax(wmax, w[i]);
    const double floor_ = fmax(reg, 1e-12 * fabs(wmax));
    for (int i = 0; i < n; ++i) w[i] = fmax(w[i], floor_);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += V[i * n + k] * w[k] * V[j * n + k];
            cov[i * n + j] = s;
        }
}
}  // namespace pv_sampler_math

// x_new = x + chol(cov) xi;  inside = lb <= x_new <= ub;  x_eval = inside ? x_new : x  (what the objective is asked for)
// returns the number of covariance matrices whose Cholesky factorisation failed (their proposal is x itself)
static int64_t pv_sampler_propose_impl(int64_t Q, int C, int dim, const double* cov, const double* x, const double* xi,
                                       const double* lb, const double* ub, double* x_new, uint8_t* inside, double* x_eval) {
    if (dim > pv_sampler_math::MAXD) return -1;
    int64_t bad = 0;
    for (int64_t qc = 0; qc < Q * C; ++qc) {
        double low[pv_sampler_math::MAXD * pv_sampler_math::MAXD];
        const bool ok = pv_sampler_math::cholesky(dim, cov + qc * dim * dim, low);
        bad += !ok;
        bool in = true;
        for (int i = 0; i < dim; ++i) {
            double s = 0.0;
            if (ok)
                for (int j = 0; j <= i; ++j) s += low[i * dim + j] * xi[qc * dim + j];
            const double v = x[qc * dim + i] + s;
            x_new[qc * dim + i] = v;
            in = in && v >= lb[i] && v <= ub[i];
        }
        inside[qc] = in;
        for (int i = 0; i < dim; ++i) x_eval[qc * dim + i] = in ? x_new[qc * dim + i] : x[qc * dim + i];
    }
    return bad;
}

// everything after the objective call of iteration `it` (1-based): Metropolis accept, proposal adaptation, swaps
// between neighbouring temperatures (hottest pair first), ladder adaptation, trace row `it`.
static int pv_sampler_step_impl(int64_t Q, int C, int dim, int64_t it, int64_t n_trace, const pv_sampler_options* o, double* x,
                                double* llh, double* lpr, const double* x_new, const double* llh_new_in,
                                const double* lpr_new, const uint8_t* inside, const double* u_acc, const double* u_swap,
                                double* betas, double* mean_hist, double* cov_hist, double* cov_scale, double* cov,
                                double* n_acc, double* n_swap, double* n_swap_try, double* trace_x, double* trace_lp,
                                double* trace_pr) {
    if (dim > pv_sampler_math::MAXD || C > 64) return -1;
    const int S = C > 1 ? C - 1 : 1;
    const double ninf = -INFINITY;
    const double rate = pow((double)it, -o->decay_constant), scale_gain = 1.0 / pow((double)(it + 1), o->decay_constant);
    for (int64_t q = 0; q < Q; ++q) {
        // ---- Metropolis step + adaptation, chain by chain -------------------------------------------------------
        for (int c = 0; c < C; ++c) {
            const int64_t qc = q * C + c;
            const double lnew = inside[qc] ? llh_new_in[qc] : ninf;
            double log_acc = betas[qc] * (lnew - llh[qc]) + (lpr_new[qc] - lpr[qc]);
            log_acc = (log_acc != log_acc) ? ninf : fmin(log_acc, 0.0);
            const bool accept = log(u_acc[qc]) < log_acc;
            if (accept) {
                for (int i = 0; i < dim; ++i) x[qc * dim + i] = x_new[qc * dim + i];
                llh[qc] = lnew;
                lpr[qc] = lpr_new[qc];
            }
            n_acc[qc] += accept;
            if (it > o->threshold_sample) {   // Haario et al.
                double dx[pv_sampler_math::MAXD];
                for (int i = 0; i < dim; ++i) {
                    mean_hist[qc * dim + i] = (1 - rate) * mean_hist[qc * dim + i] + rate * x[qc * dim + i];
                    dx[i] = x[qc * dim + i] - mean_hist[qc * dim + i];
                }
                double* ch = cov_hist + qc * dim * dim;
                for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) ch[i * dim + j] = (1 - rate) * ch[i * dim + j] + rate * (dx[i] * dx[j]);
                cov_scale[qc] *= exp((exp(log_acc) - o->target_acceptance_rate) * scale_gain);
                double* cv = cov + qc * dim * dim;
                for (int i = 0; i < dim * dim; ++i) cv[i] = cov_scale[qc] * ch[i];
                pv_sampler_math::regularize(dim, cv, o->reg_factor);
            }
        }
        // ---- swaps, hottest pair first -----------------------------------------------------------------------------
        bool swapped[64] = {};
        for (int k = C - 1; k >= 1; --k) {
            const int64_t a = q * C + (k - 1), b = q * C + k;
            double p_swap = (betas[a] - betas[b]) * (llh[b] - llh[a]);
            if (p_swap != p_swap) p_swap = ninf;
            const bool doit = log(u_swap[q * S + (k - 1)]) < p_swap;
            n_swap_try[q * S + (k - 1)] += 1;
            n_swap[q * S + (k - 1)] += doit;
            swapped[k - 1] = doit;
            if (doit) {
                for (int i = 0; i < dim; ++i) {
                    const double tmp = x[a * dim + i];
                    x[a * dim + i] = x[b * dim + i];
                    x[b * dim + i] = tmp;
                }
                double tmp = llh[a];
                llh[a] = llh[b];
                llh[b] = tmp;
                tmp = lpr[a];
                lpr[a] = lpr[b];
                lpr[b] = tmp;
            }
        }
        // ---- ladder adaptation (Vousden et al. 2016) ----------------------------------------------------------------
        if (o->adapt_betas && C > 2) {
            const double kappa = o->nu / ((double)(it + 1) + o->nu) / o->eta;
            double temps[64];
            for (int c = 0; c < C; ++c) temps[c] = 1.0 / betas[q * C + c];
            double run = temps[0];
            double prev = temps[0];
            for (int j = 0; j < C - 2; ++j) {
                const double ds = kappa * ((double)swapped[j] - (double)swapped[j + 1]);
                const double dt = (temps[j + 1] - prev) * exp(ds);
                prev = temps[j + 1];
                run += dt;
                temps[j + 1] = run;
            }
            for (int c = 0; c < C; ++c) betas[q * C + c] = 1.0 / temps[c];
        }
        // ---- trace ----------------------------------------------------------------------------------------------------
        for (int c = 0; c < C; ++c) {
            const int64_t qc = q * C + c;
            // iteration-major rows: one contiguous block per iteration (a [Q][C][n_trace] layout would touch 3 Q C
            // cache lines 160 KB apart at every iteration - that alone was 0.3 us per chain)
            const int64_t row = (int64_t)it * Q * C + qc;
            for (int i = 0; i < dim; ++i) trace_x[row * dim + i] = x[qc * dim + i];
            trace_lp[row] = -(llh[qc] + lpr[qc]);
            trace_pr[row] = -lpr[qc];
        }
    }
    return 0;
}
