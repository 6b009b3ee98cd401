// Device-resident No-U-Turn sampler on the hierarchical (population) posterior: the consumer BASELINE config 3 names ("NUTS
// logp + grad with forward sensitivities, 8 chains x 1000 individuals") with the sampler's own state on the GPU.  Multinomial NUTS
// with the generalised U-turn criterion, trees built iteratively with O(depth) momentum checkpoints, dual-averaging step size,
// one-window diagonal mass adaptation: the algorithm of optimization/nuts.py (restated there from Betancourt 2017, Phan et al.
// 2019, Hoffman & Gelman 2014), kernel by kernel -
//     pv_nuts_begin   momenta, initial energy, the tree = the current point
//     pv_nuts_start   a doubling: direction, subtree state
//     pv_nuts_drift   half kick + drift of every growing chain, objective inputs theta = exp(.), omega = exp(.)
//       -> pv_hier_logp_grad (sensitivity kernel + epilogue) -> pv_hier_finish
//     pv_nuts_leaf    gradient chain rule, second half kick, leaf weight, progressive sampling, checkpoints, U-turn tests
//     pv_nuts_merge   subtree into tree, whole-tree U-turn test
//     pv_nuts_finish  new state, dual averaging, Welford moments / mass matrix, kept sample
// - one CTA per chain, all chains in lock step: chains whose tree is finished ride along masked (they evaluate their current
// point).  Tree vectors (positions, momenta, gradients, checkpoints) are flat [n_chains][dim] arrays in HBM in the order of the
// numpy statement, x_c = (log theta_ci [individual-major], mu_c, log omega_c); the drift / leaf kernels translate to and from the
// objective's [n_par][n_chains * n_ind] layout.  The host only reads n_chains flags per leapfrog step (is any chain still
// growing) and per doubling (is every chain done) and draws the uniforms, which travel in the kernel parameters; with the same
// seed the device loop walks the numpy loop's chains up to round-off.  Individuals are NOT sharded here (one rank): the U-turn
// tests are dot products over all coordinates at every leaf.
// The reference has no gradient-based sampler (pyPESTO's adaptive Metropolis, petab_optimization.py:158-173).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/parvar_b200.h"   // pv_nuts_state, PV_NUTS_*

#define PV_NUTS_THREADS 256

__device__ __forceinline__ double pv_nuts_block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < PV_NUTS_THREADS / 32; ++k) tot += sh[k];
    return tot;
}
__device__ __forceinline__ double pv_logaddexp(double a, double b) {
    const double m = fmax(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + log1p(exp(-fabs(a - b)));
}

struct pv_nuts_view {
    const pv_nuts_state& s;
    int c;
    int64_t dim;
    __device__ double* vec(int slot) const { return s.vec + ((int64_t)slot * s.n_chains + c) * dim; }
    __device__ double* ck(int which, int level) const {
        return s.ck + (((int64_t)which * (s.max_depth + 1) + level) * s.n_chains + c) * dim;
    }
    __device__ double& sc(int slot) const { return s.sc[slot * s.n_chains + c]; }
    __device__ int& ic(int slot) const { return s.ic[slot * s.n_chains + c]; }
};

// generalised U-turn criterion on the momentum sum of a (sub)tree: rho = r_sum - (r_left + r_right) / 2
__device__ __forceinline__ bool pv_nuts_turning(const double* im, const double* rl, const double* rr, const double* rsum_a,
                                                const double* rsum_sub, const double* rsum_add, int64_t dim, double* sh) {
    // r_sum = rsum_a - rsum_sub + rsum_add (the last two may be null: r_sum = rsum_a)
    double dl = 0.0, dr = 0.0;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        double rs = rsum_a[e];
        if (rsum_sub) rs = rs - rsum_sub[e] + rsum_add[e];
        const double rho = rs - 0.5 * (rl[e] + rr[e]);
        dl += im[e] * rl[e] * rho;
        dr += im[e] * rr[e] * rho;
    }
    dl = pv_nuts_block_sum(dl, sh);
    dr = pv_nuts_block_sum(dr, sh);
    return dl <= 0.0 || dr <= 0.0;
}

__global__ void pv_nuts_begin_kernel(const pv_nuts_state s, const double* __restrict__ z) {
    __shared__ double sh[PV_NUTS_THREADS / 32];
    const int64_t dim = s.n_ind * s.n_par + 2 * s.n_par;
    const pv_nuts_view v{s, (int)blockIdx.x, dim};
    const double *x = v.vec(PV_NUTS_X), *g = v.vec(PV_NUTS_G), *im = v.vec(PV_NUTS_IM);
    const double* zc = z + (int64_t)v.c * dim;
    double K = 0.0;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        const double r0 = zc[e] / sqrt(im[e]);
        K += 0.5 * im[e] * r0 * r0;
        v.vec(PV_NUTS_XL)[e] = v.vec(PV_NUTS_XR)[e] = v.vec(PV_NUTS_XP)[e] = x[e];
        v.vec(PV_NUTS_GL)[e] = v.vec(PV_NUTS_GR)[e] = v.vec(PV_NUTS_GP)[e] = g[e];
        v.vec(PV_NUTS_RL)[e] = v.vec(PV_NUTS_RR)[e] = v.vec(PV_NUTS_RSUM)[e] = r0;
    }
    K = pv_nuts_block_sum(K, sh);
    if (threadIdx.x == 0) {
        v.sc(PV_NUTS_H0) = -v.sc(PV_NUTS_LP) + K;
        v.sc(PV_NUTS_LPP) = v.sc(PV_NUTS_LP);
        v.sc(PV_NUTS_LOGW) = 0.0;
        v.sc(PV_NUTS_SUMACC) = 0.0;
        v.ic(PV_NUTS_DEPTH) = 0;
        v.ic(PV_NUTS_DONE) = 0;
        v.ic(PV_NUTS_DIVERGED) = 0;
        v.ic(PV_NUTS_NLEAF) = 0;
    }
}

__global__ void pv_nuts_start_kernel(const pv_nuts_state s, const pv_nuts_uniforms u) {
    const int64_t dim = s.n_ind * s.n_par + 2 * s.n_par;
    const pv_nuts_view v{s, (int)blockIdx.x, dim};
    const bool right = u.u[v.c] < 0.5;
    const double *xe = v.vec(right ? PV_NUTS_XR : PV_NUTS_XL), *re = v.vec(right ? PV_NUTS_RR : PV_NUTS_RL),
                 *ge = v.vec(right ? PV_NUTS_GR : PV_NUTS_GL);
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        v.vec(PV_NUTS_XE)[e] = xe[e];
        v.vec(PV_NUTS_RE)[e] = re[e];
        v.vec(PV_NUTS_GE)[e] = ge[e];
        v.vec(PV_NUTS_SRSUM)[e] = 0.0;
        v.vec(PV_NUTS_SX)[e] = v.vec(PV_NUTS_XP)[e];
        v.vec(PV_NUTS_SG)[e] = v.vec(PV_NUTS_GP)[e];
    }
    if (threadIdx.x == 0) {
        v.ic(PV_NUTS_RIGHT) = right ? 1 : 0;
        v.ic(PV_NUTS_LEAF) = 0;
        v.ic(PV_NUTS_STURN) = 0;
        v.ic(PV_NUTS_SDIV) = 0;
        v.ic(PV_NUTS_GROWING) = v.ic(PV_NUTS_DONE) ? 0 : 1;
        v.sc(PV_NUTS_SLOGW) = -INFINITY;
        v.sc(PV_NUTS_SLP) = v.sc(PV_NUTS_LPP);
    }
}

// coordinate e of a flat chain vector <-> the objective's arrays
__device__ __forceinline__ void pv_nuts_put_eval(const pv_nuts_state& s, int c, int64_t e, double xv) {
    const int P = s.n_par;
    const int64_t LP = s.n_ind * P;
    if (e < LP) {
        const int64_t i = e / P;
        const int p = (int)(e - i * P);
        s.theta[(int64_t)p * s.n_chains * s.n_ind + (int64_t)c * s.n_ind + i] = exp(xv);
    } else if (e < LP + P) {
        s.mu_q[c * P + (e - LP)] = xv;
    } else {
        s.omega[c * P + (e - LP - P)] = exp(xv);
    }
}

__global__ void pv_nuts_drift_kernel(const pv_nuts_state s) {
    const int64_t dim = s.n_ind * s.n_par + 2 * s.n_par;
    const pv_nuts_view v{s, (int)blockIdx.x, dim};
    const bool G = v.ic(PV_NUTS_GROWING) != 0;
    const double sgn = (v.ic(PV_NUTS_RIGHT) ? 1.0 : -1.0) * v.sc(PV_NUTS_EPS);
    const double *xe = v.vec(PV_NUTS_XE), *re = v.vec(PV_NUTS_RE), *ge = v.vec(PV_NUTS_GE), *im = v.vec(PV_NUTS_IM), *x = v.vec(PV_NUTS_X);
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        const double rh = re[e] + 0.5 * sgn * ge[e];
        const double xn = xe[e] + sgn * im[e] * rh;
        v.vec(PV_NUTS_RN)[e] = rh;
        v.vec(PV_NUTS_XN)[e] = xn;
        pv_nuts_put_eval(s, v.c, e, G ? xn : x[e]);      // a chain that is not growing evaluates its current point
    }
}

__global__ void pv_nuts_leaf_kernel(const pv_nuts_state s, const pv_nuts_uniforms u, double max_energy_error) {
    __shared__ double sh[PV_NUTS_THREADS / 32];
    __shared__ int sh_flag;
    const int P = s.n_par;
    const int64_t LP = s.n_ind * P, dim = LP + 2 * P;
    const pv_nuts_view v{s, (int)blockIdx.x, dim};
    const int c = v.c;
    const bool G = v.ic(PV_NUTS_GROWING) != 0;
    const double sgn = (v.ic(PV_NUTS_RIGHT) ? 1.0 : -1.0) * v.sc(PV_NUTS_EPS);
    const double* red = s.red + (int64_t)c * (1 + 2 * P);
    const double *xn = v.vec(PV_NUTS_XN), *im = v.vec(PV_NUTS_IM);
    double *rn = v.vec(PV_NUTS_RN), *gn = v.vec(PV_NUTS_GN);
    // gradient of the transformed density at the evaluation point (chain rule, + 1 from the Jacobian), finiteness
    int bad_local = !isfinite(red[0]);
    double J = 0.0;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        double g;
        if (e < LP) {
            const int64_t i = e / P;
            const int p = (int)(e - i * P);
            const int64_t k = (int64_t)p * s.n_chains * s.n_ind + (int64_t)c * s.n_ind + i;
            g = fma(s.dtheta[k], s.theta[k], 1.0);
            J += xn[e];
        } else if (e < LP + P) {
            g = red[1 + (e - LP)];
        } else {
            g = fma(red[1 + (e - LP)], s.omega[c * P + (e - LP - P)], 1.0);
            J += xn[e];
        }
        bad_local |= !isfinite(g);
        gn[e] = g;
    }
    if (threadIdx.x == 0) sh_flag = 0;
    __syncthreads();
    if (bad_local) sh_flag = 1;
    __syncthreads();
    const bool bad = sh_flag != 0;
    double K = 0.0;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        const double g = bad ? 0.0 : gn[e];
        const double r = rn[e] + 0.5 * sgn * g;       // rn held the half-kicked momentum
        gn[e] = g;
        rn[e] = r;
        K += 0.5 * im[e] * r * r;
    }
    K = pv_nuts_block_sum(K, sh);
    J = pv_nuts_block_sum(J, sh);
    const double lpn = red[0] + J;
    double dH = bad ? INFINITY : (-lpn + K) - v.sc(PV_NUTS_H0);
    if (dH != dH) dH = INFINITY;
    const bool leaf_div = dH > max_energy_error;
    const double lw = -dH;
    const double new_log_w = pv_logaddexp(v.sc(PV_NUTS_SLOGW), lw);
    double p_take = exp(lw - new_log_w);
    if (p_take != p_take) p_take = 0.0;
    const bool take = G && !leaf_div && (u.u[c] < p_take);
    const int leaf = v.ic(PV_NUTS_LEAF), depth = v.ic(PV_NUTS_DEPTH);
    double *srs = v.vec(PV_NUTS_SRSUM);
    if (G) {
        for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
            srs[e] += rn[e];
            if (take) {
                v.vec(PV_NUTS_SX)[e] = xn[e];
                v.vec(PV_NUTS_SG)[e] = gn[e];
            }
        }
    }
    __syncthreads();
    // U-turn inside the subtree via momentum checkpoints (iterative NUTS)
    bool turning = false;
    if (G) {
        const int idx_max = __popc(leaf >> 1);
        const int n_tr = __ffs(~leaf) - 1;                 // trailing ones of leaf
        const int idx_min = idx_max - n_tr + 1;
        if ((leaf & 1) == 0) {
            double *rc = v.ck(0, idx_max), *rsc = v.ck(1, idx_max);
            for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
                rc[e] = rn[e];
                rsc[e] = srs[e];
            }
        } else {
            for (int i = idx_max; i >= idx_min && !turning; --i)
                turning = pv_nuts_turning(im, v.ck(0, i), rn, srs, v.ck(1, i), v.ck(0, i), dim, sh);
        }
    }
    __syncthreads();
    if (G) {
        for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
            v.vec(PV_NUTS_XE)[e] = xn[e];
            v.vec(PV_NUTS_RE)[e] = rn[e];
            v.vec(PV_NUTS_GE)[e] = gn[e];
        }
    }
    if (threadIdx.x == 0 && G) {
        v.sc(PV_NUTS_SUMACC) += fmin(1.0, exp(fmin(-dH, 0.0)));
        v.ic(PV_NUTS_NLEAF) += 1;
        if (take) v.sc(PV_NUTS_SLP) = lpn;
        v.sc(PV_NUTS_SLOGW) = new_log_w;
        const int sturn = v.ic(PV_NUTS_STURN) | (turning ? 1 : 0), sdiv = v.ic(PV_NUTS_SDIV) | (leaf_div ? 1 : 0);
        v.ic(PV_NUTS_STURN) = sturn;
        v.ic(PV_NUTS_SDIV) = sdiv;
        v.ic(PV_NUTS_LEAF) = leaf + 1;
        v.ic(PV_NUTS_GROWING) = (!sturn && !sdiv && leaf + 1 < (1 << depth)) ? 1 : 0;
    }
    if (threadIdx.x == 0 && s.host_flags) s.host_flags[c] = v.ic(PV_NUTS_GROWING);
}

__global__ void pv_nuts_merge_kernel(const pv_nuts_state s, const pv_nuts_uniforms u) {
    __shared__ double sh[PV_NUTS_THREADS / 32];
    const int64_t dim = s.n_ind * s.n_par + 2 * s.n_par;
    const pv_nuts_view v{s, (int)blockIdx.x, dim};
    const int c = v.c;
    const bool act = v.ic(PV_NUTS_DONE) == 0;
    const bool sturn = v.ic(PV_NUTS_STURN) != 0, sdiv = v.ic(PV_NUTS_SDIV) != 0, right = v.ic(PV_NUTS_RIGHT) != 0;
    const bool ok = act && !sturn && !sdiv;
    const double slw = v.sc(PV_NUTS_SLOGW), lw = v.sc(PV_NUTS_LOGW);
    const bool swap = ok && (u.u[c] < exp(fmin(slw - lw, 0.0)));
    if (ok) {
        double *xend = v.vec(right ? PV_NUTS_XR : PV_NUTS_XL), *rend = v.vec(right ? PV_NUTS_RR : PV_NUTS_RL),
               *gend = v.vec(right ? PV_NUTS_GR : PV_NUTS_GL);
        for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
            if (swap) {
                v.vec(PV_NUTS_XP)[e] = v.vec(PV_NUTS_SX)[e];
                v.vec(PV_NUTS_GP)[e] = v.vec(PV_NUTS_SG)[e];
            }
            v.vec(PV_NUTS_RSUM)[e] += v.vec(PV_NUTS_SRSUM)[e];
            xend[e] = v.vec(PV_NUTS_XE)[e];
            rend[e] = v.vec(PV_NUTS_RE)[e];
            gend[e] = v.vec(PV_NUTS_GE)[e];
        }
    }
    __syncthreads();
    const bool whole_turn = pv_nuts_turning(v.vec(PV_NUTS_IM), v.vec(PV_NUTS_RL), v.vec(PV_NUTS_RR), v.vec(PV_NUTS_RSUM), nullptr,
                                            nullptr, dim, sh);
    if (threadIdx.x == 0 && act) {
        if (swap) v.sc(PV_NUTS_LPP) = v.sc(PV_NUTS_SLP);
        if (ok) v.sc(PV_NUTS_LOGW) = pv_logaddexp(lw, slw);
        const int depth = v.ic(PV_NUTS_DEPTH) + 1;
        v.ic(PV_NUTS_DEPTH) = depth;
        if (sdiv) v.ic(PV_NUTS_DIVERGED) = 1;
        if (sturn || sdiv || whole_turn || depth >= s.max_depth) v.ic(PV_NUTS_DONE) = 1;
    }
    if (threadIdx.x == 0 && s.host_flags) s.host_flags[s.n_chains + c] = v.ic(PV_NUTS_DONE);
}

// mode 0: adopt the evaluation of the start point (gn / red of the start -> state); 1: warm-up iteration; 2: sampling iteration
__global__ void pv_nuts_finish_kernel(const pv_nuts_state s, int mode, int64_t it, int64_t n_warmup, double target_accept,
                                      int adapt_mass, int64_t w_n) {
    __shared__ double sh[PV_NUTS_THREADS / 32];
    const int P = s.n_par;
    const int64_t LP = s.n_ind * P, dim = LP + 2 * P;
    const pv_nuts_view v{s, (int)blockIdx.x, dim};
    const int c = v.c;
    double *x = v.vec(PV_NUTS_X), *g = v.vec(PV_NUTS_G), *im = v.vec(PV_NUTS_IM);
    if (mode == 0) {
        // start: the leaf kernel left the transformed gradient of the start point in GN and its Jacobian-free logp in red
        double J = 0.0;
        for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
            g[e] = v.vec(PV_NUTS_GN)[e];
            if (e < LP || e >= LP + P) J += x[e];
        }
        J = pv_nuts_block_sum(J, sh);
        if (threadIdx.x == 0) v.sc(PV_NUTS_LP) = s.red[(int64_t)c * (1 + 2 * P)] + J;
        return;
    }
    const bool welford = mode == 1 && adapt_mass && w_n > 0;
    const bool finish_mass = mode == 1 && it == n_warmup - 1 && adapt_mass && w_n > 10;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        const double xv = v.vec(PV_NUTS_XP)[e];
        x[e] = xv;
        g[e] = v.vec(PV_NUTS_GP)[e];
        if (welford) {
            double* wm = v.vec(PV_NUTS_WMEAN);
            const double d = xv - wm[e], m = wm[e] + d / (double)w_n;
            wm[e] = m;
            v.vec(PV_NUTS_WM2)[e] += d * (xv - m);
        }
        if (finish_mass) {
            const double var = v.vec(PV_NUTS_WM2)[e] / (double)(w_n - 1);
            im[e] = ((double)w_n / ((double)w_n + 5.0)) * var + 1e-3 * (5.0 / ((double)w_n + 5.0));
        }
        if (mode == 2) s.samples[((int64_t)c * s.n_samples + (it - n_warmup)) * dim + e] = xv;
    }
    if (threadIdx.x == 0) {
        v.sc(PV_NUTS_LP) = v.sc(PV_NUTS_LPP);
        const int n_leaf = v.ic(PV_NUTS_NLEAF);
        const double a = v.sc(PV_NUTS_SUMACC) / (double)(n_leaf > 1 ? n_leaf : 1);
        if (mode == 1) {
            constexpr double gamma = 0.05, t0 = 10.0, kappa = 0.75;
            const double m = (double)(it + 1);
            const double h_bar = (1.0 - 1.0 / (m + t0)) * v.sc(PV_NUTS_DA_HBAR) + (target_accept - a) / (m + t0);
            const double log_eps = v.sc(PV_NUTS_DA_MU) - sqrt(m) / gamma * h_bar;
            const double eta = pow(m, -kappa);
            const double leb = eta * log_eps + (1.0 - eta) * v.sc(PV_NUTS_DA_LEB);
            v.sc(PV_NUTS_DA_HBAR) = h_bar;
            v.sc(PV_NUTS_DA_LEB) = leb;
            v.sc(PV_NUTS_EPS) = (it == n_warmup - 1) ? exp(leb) : exp(log_eps);
        } else {
            const int64_t k = (int64_t)c * s.n_samples + (it - n_warmup);
            s.samples_lp[k] = v.sc(PV_NUTS_LPP);
            s.samples_depth[k] = v.ic(PV_NUTS_DEPTH);
            v.sc(PV_NUTS_ACCSUM) += a;
            v.ic(PV_NUTS_NDIV) += v.ic(PV_NUTS_DIVERGED);
        }
    }
}
