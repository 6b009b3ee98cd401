// Device-resident batched multistart optimiser: the loop of optimization/optimizer.py::minimize (damped Gauss-Newton /
// Levenberg-Marquardt on the Fisher information + prior precision, box projection, gain-ratio damping update) for Q
// problems x S starts at once, with nothing of an iteration on the host.  Reference: pyPESTO's multistart
// `pypesto.optimize.minimize` with FidesOptimizer, 100 starts per problem (src/parvar/optimization/petab_optimization.py:
// 99-156), whose every iteration is one AMICI call returning fval, gradient and the FIM; here one iteration is
//     pv_opt_propose_kernel  ->  the batched objective (pv_logp_grad_fim: S*Q*G trajectories)  ->  pv_opt_update_kernel
// for all starts of all problems.  The host only polls the number of active starts every few iterations.
// Same arithmetic and the same accept / convergence rules as the numpy statement (tests compare the two).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define PV_OPT_MAXD 16

struct pv_opt_dev {
    int64_t Q;                  // problems
    int S, G, NP, dim, NS;      // starts, groups, parameters per group, NP * G, compiled sensitivity parameters of the model
    double *x, *f, *g, *H, *lam;            // [Q][S][dim], [Q][S], [Q][S][dim], [Q][S][dim][dim], [Q][S]
    unsigned char *active, *converged;      // [Q][S]
    int* n_iter;                            // [Q][S]
    double *x_try, *p, *pred;               // [Q][S][dim] x 2, [Q][S]
    double* P;                              // [NP][S*Q*G]  objective input: trajectory (s*Q + q)*G + g <- x_eval[q][s][p*G + g]
    const double *logp, *grad, *fim;        // objective output: [B], [NS][B], [NS(NS+1)/2][B]
    const double *lb, *ub;                  // [dim]
    const double *prior_loc, *prior_scale;  // [Q][dim], NaN loc = none
    unsigned long long* n_evals;            // accumulated number of objective evaluations of active starts
    int* n_active;                          // active starts after the latest update
    int sens_row[8];                        // column p -> row of grad / index into the FIM triangle
    double fatol, frtol, gtol;
};

// upper-triangle index of (s, r), s <= r, row-major ((0,0),(0,1),..,(1,1),..)
__device__ __forceinline__ int pv_opt_tri(int s, int r, int ns) { return s * ns - s * (s - 1) / 2 + (r - s); }

// one thread per (problem, start): LM step from (g, H, lam), projection onto the box, predicted decrease, objective input
__global__ void pv_opt_propose_kernel(const pv_opt_dev o, int init) {
    const int64_t qs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qs >= o.Q * o.S) return;
    const int dim = o.dim;
    const int64_t q = qs / o.S;
    const int s = (int)(qs % o.S);
    const double* x = o.x + qs * dim;
    double* xt = o.x_try + qs * dim;
    const bool act = init ? true : o.active[qs] != 0;
    if (init) {
        for (int i = 0; i < dim; ++i) xt[i] = x[i];
    } else {
        double A[PV_OPT_MAXD * PV_OPT_MAXD], Hs[PV_OPT_MAXD * PV_OPT_MAXD], b[PV_OPT_MAXD], gs[PV_OPT_MAXD], D[PV_OPT_MAXD];
        const double lam = o.lam[qs];
        for (int i = 0; i < dim; ++i) {
            const double gi = o.g[qs * dim + i];
            gs[i] = isfinite(gi) ? gi : 0.0;
            for (int j = 0; j < dim; ++j) {
                const double h = o.H[(qs * dim + i) * dim + j];
                Hs[i * dim + j] = isfinite(h) ? h : 0.0;
            }
        }
        for (int i = 0; i < dim; ++i) D[i] = fmax(Hs[i * dim + i], 1e-12);
        for (int i = 0; i < dim; ++i) {
            for (int j = 0; j < dim; ++j) A[i * dim + j] = Hs[i * dim + j];
            A[i * dim + i] += lam * D[i];
            b[i] = -gs[i];
        }
        // Gaussian elimination with partial pivoting (what numpy.linalg.solve does); a singular system falls back to the
        // scaled gradient step
        bool singular = false;
        for (int k = 0; k < dim && !singular; ++k) {
            int piv = k;
            double best = fabs(A[k * dim + k]);
            for (int i = k + 1; i < dim; ++i)
                if (fabs(A[i * dim + k]) > best) {
                    best = fabs(A[i * dim + k]);
                    piv = i;
                }
            if (!(best > 0.0)) {
                singular = true;
                break;
            }
            if (piv != k) {
                for (int j = 0; j < dim; ++j) {
                    const double tmp = A[k * dim + j];
                    A[k * dim + j] = A[piv * dim + j];
                    A[piv * dim + j] = tmp;
                }
                const double tmp = b[k];
                b[k] = b[piv];
                b[piv] = tmp;
            }
            for (int i = k + 1; i < dim; ++i) {
                const double m = A[i * dim + k] / A[k * dim + k];
                for (int j = k; j < dim; ++j) A[i * dim + j] -= m * A[k * dim + j];
                b[i] -= m * b[k];
            }
        }
        double pstep[PV_OPT_MAXD];
        if (!singular) {
            for (int i = dim - 1; i >= 0; --i) {
                double acc = b[i];
                for (int j = i + 1; j < dim; ++j) acc -= A[i * dim + j] * pstep[j];
                pstep[i] = acc / A[i * dim + i];
            }
        } else {
            for (int i = 0; i < dim; ++i) pstep[i] = -gs[i] / (D[i] * (1.0 + lam));
        }
        double gp = 0.0, php = 0.0;
        for (int i = 0; i < dim; ++i) {
            const double v = fmin(fmax(x[i] + pstep[i], o.lb[i]), o.ub[i]);
            xt[i] = v;
            pstep[i] = v - x[i];
            o.p[qs * dim + i] = pstep[i];
        }
        for (int i = 0; i < dim; ++i) {
            gp += gs[i] * pstep[i];
            double hp = 0.0;
            for (int j = 0; j < dim; ++j) hp += Hs[i * dim + j] * pstep[j];
            php += pstep[i] * hp;
        }
        o.pred[qs] = -(gp + 0.5 * php);
    }
    const int64_t BT = (int64_t)o.S * o.Q * o.G;
    for (int i = 0; i < dim; ++i) {
        const int pp = i / o.G, gg = i % o.G;
        o.P[(int64_t)pp * BT + ((int64_t)s * o.Q + q) * o.G + gg] = act ? xt[i] : x[i];
    }
}

// one thread per (problem, start): assemble (f, g, H) of the evaluated point, accept / reject, damping, convergence
__global__ void pv_opt_update_kernel(const pv_opt_dev o, int init) {
    constexpr double LOG_2PI = 1.8378770664093454835606594728112;
    const int64_t qs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qs >= o.Q * o.S) return;
    const int dim = o.dim, G = o.G, NP = o.NP, NS = o.NS;
    const int64_t q = qs / o.S;
    const int s = (int)(qs % o.S);
    const int64_t BT = (int64_t)o.S * o.Q * G;
    const bool act = init ? true : o.active[qs] != 0;
    if (!act) return;
    const double* xe = o.x_try + qs * dim;
    // ---- objective at x_try: negative log posterior, gradient, FIM + prior precision ---------------------------------
    double llh = 0.0, gt[PV_OPT_MAXD], Ht[PV_OPT_MAXD * PV_OPT_MAXD];
    for (int i = 0; i < dim * dim; ++i) Ht[i] = 0.0;
    for (int gg = 0; gg < G; ++gg) {
        const int64_t tr = ((int64_t)s * o.Q + q) * G + gg;
        llh += o.logp[tr];
        for (int pa = 0; pa < NP; ++pa) {
            const int sa = o.sens_row[pa];
            gt[pa * G + gg] = -o.grad[(int64_t)sa * BT + tr];
            for (int pb = 0; pb < NP; ++pb) {
                const int sb = o.sens_row[pb];
                const int lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
                Ht[(pa * G + gg) * dim + (pb * G + gg)] = o.fim[(int64_t)pv_opt_tri(lo, hi, NS) * BT + tr];
            }
        }
    }
    const bool bad = !isfinite(llh);
    double lpr = 0.0;
    for (int i = 0; i < dim; ++i) {
        const double loc = o.prior_loc[q * dim + i], sc = o.prior_scale[q * dim + i];
        if (loc == loc) {
            const double z = (xe[i] - loc) / sc;
            lpr += -0.5 * z * z - 0.5 * (LOG_2PI + 2.0 * log(sc));
            gt[i] += z / sc;
            Ht[i * dim + i] += 1.0 / (sc * sc);
        }
    }
    double f_try = bad ? INFINITY : -(llh + lpr);
    if (bad)
        for (int i = 0; i < dim; ++i) gt[i] = __longlong_as_double(0x7ff8000000000000LL);
    if (init) {
        o.f[qs] = f_try;
        for (int i = 0; i < dim; ++i) o.g[qs * dim + i] = gt[i];
        for (int i = 0; i < dim * dim; ++i) o.H[qs * dim * dim + i] = Ht[i];
        o.lam[qs] = 1e-2;
        o.n_iter[qs] = 0;
        o.converged[qs] = 0;
        const bool a = isfinite(f_try);
        o.active[qs] = a;
        if (a) atomicAdd(o.n_active, 1);
        atomicAdd(o.n_evals, 1ull);
        return;
    }
    atomicAdd(o.n_evals, 1ull);
    if (!isfinite(f_try)) f_try = INFINITY;
    const double f = o.f[qs];
    const double rho = (f - f_try) / fmax(o.pred[qs], 1e-300);
    const bool take = (f_try < f) && (rho > 1e-4);
    const double df = take ? f - f_try : INFINITY;
    double* x = o.x + qs * dim;
    double lam = o.lam[qs];
    if (take) {
        for (int i = 0; i < dim; ++i) {
            x[i] = xe[i];
            o.g[qs * dim + i] = gt[i];
        }
        for (int i = 0; i < dim * dim; ++i) o.H[qs * dim * dim + i] = Ht[i];
        o.f[qs] = f_try;
        lam = fmax(lam * (rho > 0.75 ? 1.0 / 3.0 : 1.0), 1e-12);
    } else {
        lam = fmin(lam * 4.0, 1e12);
    }
    o.lam[qs] = lam;
    o.n_iter[qs] += 1;
    const double fnow = take ? f_try : f;
    double gmax = 0.0, pmax = 0.0, xmax = 0.0;
    for (int i = 0; i < dim; ++i) {
        const double gi = o.g[qs * dim + i];
        const bool pinned = (x[i] <= o.lb[i] && gi > 0.0) || (x[i] >= o.ub[i] && gi < 0.0);
        const double gpi = pinned ? 0.0 : fabs(gi);
        gmax = (gpi > gmax || gpi != gpi) ? gpi : gmax;          // a NaN gradient never counts as converged (numpy: NaN < gtol is False)
        pmax = fmax(pmax, fabs(o.p[qs * dim + i]));
        xmax = fmax(xmax, fabs(x[i]));
    }
    const bool done = (df < o.fatol + o.frtol * fabs(fnow)) || (gmax < o.gtol) || (lam >= 1e12) || (pmax < 1e-14 * (1.0 + xmax));
    if (done) {
        if (isfinite(fnow)) o.converged[qs] = 1;
        o.active[qs] = 0;
    } else {
        atomicAdd(o.n_active, 1);
    }
}
