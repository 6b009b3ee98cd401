// RODAS4 with forward sensitivities, ONE WARP PER BLOCK of the augmented system (the stiff gradient path of
// north_star subsystems 2-4: logp + gradient (+ Fisher information) of the PBPK model, reference call
// src/parvar/optimization/petab_optimization.py:144-151 -> AMICI CVODES with forward sensitivities).
//
// Why not one trajectory per lane (pv_rodas4_lane<M, true>): the augmented state of icg_body_flat is 6 stage
// vectors x 36 + z + stage point + LU (31) + hoisted coefficients (32 + 23) = 374 doubles against 127 in 255
// registers.  That kernel keeps 2.7 KB per lane in local memory, its 102 MB footprint falls out of the L2 (81 GB of
// DRAM writes per 128 000-trajectory launch) and its 117 KB step body misses the instruction cache
// (profiles/r01_icg_c4_logp_grad_rodas4_ncu_full.txt).
//
// Here a CTA of (1 + NS) warps integrates 32 trajectories: lane l of warp 0 owns the state block of trajectory l,
// lane l of warp k its k-th sensitivity block.  The augmented Jacobian is block lower triangular, [[J,0],[H_k,J]], so
//   * every block's stage vectors K1..K5 live in shared memory, [block][stage][state][32] lane-minor (conflict free),
//     indexed by the RUN-TIME stage: the stage loop is a real loop (its body - RHS, H-term, LU solve - exists once and
//     stays in the instruction cache) and the stage sums are short loops over a coefficient table in constant memory.
//     (First version: K1..K5 in registers, selected by a switch on the stage - 23 % of its instructions were the
//     register moves of that switch and 10 % spills, profiles/r01_icg_c4_logp_grad_split_ncu_full.txt.)
//   * what the blocks share - hoisted coefficients c, dc, the LU of W = I/(gamma h) - J, the state at the start of the
//     step - lives in shared memory too, written once per trajectory / step,
//   * warp 0 runs one stage AHEAD: in tick i it computes stage i of the state block and publishes the stage point
//     Y_i (double buffered) and K_i^y (its own stage-vector slot), while the sensitivity warps compute stage i-1 from
//     what it published in the previous tick; one __syncthreads per tick,
//   * every warp executes only its own role's code: no intra-warp role divergence and a third of the code per warp,
//   * the step-size controller runs replicated in all warps from the same shared error partials, so all warps take
//     identical accept / reject / output decisions without further communication.
// The likelihood lives in warp 0 (it owns the observables), the gradient component k in warp k: at a grid point warp 0
// publishes d lp / d f per observable, warp k folds it with its own d f / d theta_k.
#pragma once
#include "pv_kernels.cuh"

#ifndef PV_MINB_SPLIT
#define PV_MINB_SPLIT 2
#endif

// RODAS4 stage coefficients by run-time stage (row st - 1, column j = index of K_{j+1}); stage 6 evaluates at the embedded
// solution z + sum_j A5j K_j + K5, hence the 1 in its A row
__constant__ double pv_rodas4_A[6][5] = {
    {0, 0, 0, 0, 0},
    {pv_rodas4::A21, 0, 0, 0, 0},
    {pv_rodas4::A31, pv_rodas4::A32, 0, 0, 0},
    {pv_rodas4::A41, pv_rodas4::A42, pv_rodas4::A43, 0, 0},
    {pv_rodas4::A51, pv_rodas4::A52, pv_rodas4::A53, pv_rodas4::A54, 0},
    {pv_rodas4::A51, pv_rodas4::A52, pv_rodas4::A53, pv_rodas4::A54, 1.0}};
__constant__ double pv_rodas4_C[6][5] = {
    {0, 0, 0, 0, 0},
    {pv_rodas4::C21, 0, 0, 0, 0},
    {pv_rodas4::C31, pv_rodas4::C32, 0, 0, 0},
    {pv_rodas4::C41, pv_rodas4::C42, pv_rodas4::C43, 0, 0},
    {pv_rodas4::C51, pv_rodas4::C52, pv_rodas4::C53, pv_rodas4::C54, 0},
    {pv_rodas4::C61, pv_rodas4::C62, pv_rodas4::C63, pv_rodas4::C64, pv_rodas4::C65}};

// element i of this lane's trajectory in a lane-minor shared-memory array [n][32]
struct pv_smem_vec {
    double* p;   // &array[0][lane]
    PV_D double operator[](int i) const { return p[i * 32]; }
    PV_D void set(int i, double v) const { p[i * 32] = v; }
};

template <class M>
struct pv_split_layout {
    static constexpr int NB = 1 + M::NS;
    // doubles per CTA (32 trajectories), in this order
    static constexpr int C = 0;
    static constexpr int DC = C + M::NC * 32;
    static constexpr int W = DC + M::NDC * 32;
    static constexpr int YN = W + M::NLU * 32;
    static constexpr int YS = YN + M::NX * 32;            // [2][NX][32] stage point of the state block
    static constexpr int KS = YS + 2 * M::NX * 32;        // [NB][5][NX][32] stage vectors K1..K5 of every block
    static constexpr int Z0 = KS;                         // [NB][NX][32] initial blocks (set-up hand-over; aliases KS)
    static constexpr int K6Y = KS + NB * 5 * M::NX * 32;  // [NX][32] K6 of the state block
    static constexpr int ERR = K6Y + M::NX * 32;          // [2][NB][32] error-norm / initial-step partials
    static constexpr int ZERO = ERR + 2 * NB * 32;        // [NX][32] zeros: the "K_1" stage 1 sums over (see the stage sums)
    static constexpr int FIXED = ZERO + M::NX * 32;
    // then, sized at launch: OW [n_obs][32], OWI [n_obs][32], DF [NS][n_obs][32], tgrid [T], theta [NTH], y0 [NX]
    static size_t bytes(int T, int n_obs) {
        return sizeof(double) * ((size_t)FIXED + (size_t)(2 + M::NS) * n_obs * 32 + T + M::NTH + M::NX);
    }
};

template <class M, int MODE>
__global__ void __launch_bounds__(32 * (1 + M::NS), PV_MINB_SPLIT)
pv_rodas4_split_kernel(const __grid_constant__ pv_launch_params_m<M> p) {
    using namespace pv_rodas4;
    using L = pv_split_layout<M>;
    constexpr int NX = M::NX, NS = M::NS, NB = 1 + NS, N = NX * NB;
    constexpr bool FIM = (MODE & PV_MODE_FIM) != 0;
    static_assert((MODE & PV_MODE_SENS) && (MODE & PV_MODE_LOGP) && !(MODE & PV_MODE_TRAJ), "logp + gradient modes only");
    static_assert(NS >= 1 && NS <= 2, "one warp per block: instantiated for 1 or 2 sensitivity parameters");
    extern __shared__ __align__(16) double sh[];
    __shared__ long long s_item;
    const int T = p.T, n_obs = p.n_obs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sh_ow = sh + L::FIXED;                 // d lp / d f * obs scale            [n_obs][32]
    double* sh_owi = sh_ow + n_obs * 32;           // Fisher weight * obs scale^2       [n_obs][32]
    double* sh_df = sh_owi + n_obs * 32;           // sensitivity of the observables    [NS][n_obs][32]
    double* sh_t = sh_df + NS * n_obs * 32;        // [T]
    double* sh_theta = sh_t + T;                   // [NTH]
    double* sh_y0 = sh_theta + M::NTH;             // [NX]
    for (int i = threadIdx.x; i < T; i += blockDim.x) sh_t[i] = p.tgrid[i];
    for (int i = threadIdx.x; i < M::NTH; i += blockDim.x) sh_theta[i] = p.theta_base[i];
    for (int i = threadIdx.x; i < NX; i += blockDim.x) sh_y0[i] = p.y0_base[i];
    for (int i = threadIdx.x; i < NX * 32; i += blockDim.x) sh[L::ZERO + i] = 0.0;
    const pv_smem_vec C{sh + L::C + lane}, DC{sh + L::DC + lane}, Wsm{sh + L::W + lane}, YN{sh + L::YN + lane};
    const pv_smem_vec Z0{sh + L::Z0 + lane};
    double* const err_sh = sh + L::ERR + lane;     // [2][NB] partials of this lane's trajectory, stride 32
    const int64_t n_items = p.n_items_dev ? (int64_t)*p.n_items_dev : p.B;
    const double tend = p.tgrid[T - 1];
    const float inv_n = 1.0f / N;

    for (;;) {
        __syncthreads();   // previous batch completely finished (s_item, shared arrays)
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(p.counter, 32ull);
        __syncthreads();
        if (s_item >= n_items) break;
        // batches are handed out from the END of the work list (the high-key end under the locality schedule: longest first)
        const int64_t item0 = (n_items - 1) / 32 * 32 - s_item;
        const int64_t item = item0 + lane;
        const bool valid = item < n_items;
        const int64_t b = valid ? (p.index_map ? (int64_t)p.index_map[item] : item) : 0;

        // ---- set-up by warp 0: theta, hoisted coefficients, initial blocks -------------------------------------
        double obs[NX];          // observable scale (warp 0)
        if (warp == 0 && valid) {
            double pc[PV_MAX_COLS];
#pragma unroll
            for (int j = 0; j < PV_MAX_COLS; ++j) pc[j] = (j < p.n_cols) ? __ldg(p.P + (int64_t)j * p.B + b) : 0.0;
            double th[M::NTH];
#pragma unroll
            for (int i = 0; i < M::NTH; ++i) {
                double v = sh_theta[i];
#pragma unroll
                for (int j = 0; j < PV_MAX_COLS; ++j) v = (j < p.n_cols && p.col_target[j] == i) ? pc[j] : v;
                th[i] = v;
            }
            double c[M::NC], y0[NX], dc[M::NDC], dy0[M::NSD][NX];
            M::hoist(th, c, y0, obs);
            M::hoist_sens(th, dc, dy0);
#pragma unroll
            for (int i = 0; i < M::NC; ++i) C.set(i, c[i]);
#pragma unroll
            for (int i = 0; i < M::NDC; ++i) DC.set(i, dc[i]);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double v = sh_y0[i];
                bool over = (v == v);
                v = over ? v : y0[i];
#pragma unroll
                for (int j = 0; j < PV_MAX_COLS; ++j)
                    if (j < p.n_cols && p.col_target[j] == M::NTH + i) {
                        v = pc[j];
                        over = true;
                    }
                Z0.set(i, v);
#pragma unroll
                for (int s = 0; s < NS; ++s) Z0.set(NX * (1 + s) + i, over ? 0.0 : dy0[s][i]);
            }
        }
        __syncthreads();

        // ---- per-lane state: the control variables are replicated in every warp and evolve identically ----------
        double z[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) z[i] = valid ? Z0[warp * NX + i] : 0.0;
        double* const kmine = sh + L::KS + (warp * 5 * NX) * 32 + lane;   // this block's K_{j+1}[i] = kmine[(j * NX + i) * 32]
        double t = p.t0, h = 0.0, h_acc = 0.0;
        float err_acc = 1e-2f;
        bool last_rejected = false, running = valid;
        int jout = 0, nstep = 0, n_acc = 0, n_rej = 0, rc = PV_OK;
        double lp = 0.0;                       // warp 0
        double g = 0.0;                        // warp k: d logp / d theta_k
        double fim[M::NSD] = {};               // warp k: Fisher information row k (entries (k, r >= k))
        const double* qrow = p.stats + (b % p.n_data) * (int64_t)(T * n_obs * PV_QSTRIDE);   // warp 0

        // one grid point: zo = this warp's block at tgrid[jout]; folds it into logp (warp 0) / gradient (warp k).
        // CTA-uniform: every thread calls it the same number of times (`act` masks the lanes that have no output).
        auto sink = [&](bool act, const double (&zo)[NX]) {
            if (warp == 0) {
                if (act) {
                    for (int k = 0; k < n_obs; ++k) {
                        const int idx = p.obs_idx[k];
                        const double sc = pv_pick<NX>(obs, 0, idx);
                        const double f = pv_pick<NX>(zo, 0, idx) * sc;
                        const double q0 = qrow[k * PV_QSTRIDE], q1 = qrow[k * PV_QSTRIDE + 1], q2 = qrow[k * PV_QSTRIDE + 2];
                        double w = 0.0, wi = 0.0;
                        if (p.noise_model == 0) {
                            lp += fma(f, fma(q2, f, q1), q0);
                            w = fma(2.0 * q2, f, q1);
                            wi = -2.0 * q2;
                        } else if (q2 != 0.0) {
                            const double lf = log(f), rf = 1.0 / f;
                            lp += fma(lf, fma(q2, lf, q1), q0);
                            w = fma(2.0 * q2, lf, q1) * rf;
                            wi = -2.0 * q2 * rf * rf;
                        }
                        sh_ow[k * 32 + lane] = w * sc;
                        if constexpr (FIM) sh_owi[k * 32 + lane] = wi * sc * sc;
                    }
                    qrow += n_obs * PV_QSTRIDE;
                }
            } else if constexpr (FIM) {
                if (act)
                    for (int k = 0; k < n_obs; ++k) sh_df[((warp - 1) * n_obs + k) * 32 + lane] = pv_pick<NX>(zo, 0, p.obs_idx[k]);
            }
            __syncthreads();
            if (warp > 0 && act) {
                for (int k = 0; k < n_obs; ++k) {
                    const double df = pv_pick<NX>(zo, 0, p.obs_idx[k]);
                    g = fma(sh_ow[k * 32 + lane], df, g);
                    if constexpr (FIM) {
                        const double wdf = sh_owi[k * 32 + lane] * df;
#pragma unroll
                        for (int r = 0; r < NS; ++r)
                            if (r >= warp - 1) fim[r] = fma(wdf, sh_df[(r * n_obs + k) * 32 + lane], fim[r]);
                    }
                }
            }
            __syncthreads();   // the exchange buffers may be rewritten
        };

        // ---- grid points at t0 (CTA-uniform trip count: t0 and the grid are launch-wide) -------------------------
        while (jout < T && sh_t[jout] <= t) {
            sink(running, z);
            ++jout;
        }
        if (jout >= T) running = false;

        // ---- initial step size from all blocks (Hairer's heuristic, first part) ----------------------------------
        {
            if (warp == 0 && running) {
#pragma unroll
                for (int i = 0; i < NX; ++i) YN.set(i, z[i]);
            }
            __syncthreads();
            float d0 = 0.f, d1 = 0.f;
            if (running) {
                double F0[NX];
                if (warp == 0) {
                    M::rhs(t, z, C, F0);
                } else {
                    if (warp == 1) M::template sens_rhs_block<0>(t, YN, z, C, DC, F0);
                    if constexpr (NS > 1)
                        if (warp == 2) M::template sens_rhs_block<(NS > 1 ? 1 : 0)>(t, YN, z, C, DC, F0);
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const float sc = (float)(p.atol + p.rtol * fabs(z[i]));
                    const float a = (float)z[i] / sc, q = (float)F0[i] / sc;
                    d0 += a * a;
                    d1 += q * q;
                }
            }
            err_sh[warp * 32] = (double)d0;
            err_sh[(NB + warp) * 32] = (double)d1;
            __syncthreads();
            d0 = 0.f;
            d1 = 0.f;
#pragma unroll
            for (int w = 0; w < NB; ++w) {
                d0 += (float)err_sh[w * 32];
                d1 += (float)err_sh[(NB + w) * 32];
            }
            d0 = sqrtf(d0 * inv_n);
            d1 = sqrtf(d1 * inv_n);
            h = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6 : 0.01 * (double)(d0 / d1);
            h = fmin(fmax(h, 1e-8), tend - t);
            h_acc = h;
        }

        // ---- steps --------------------------------------------------------------------------------------------
        while (__syncthreads_or(running)) {
            bool last = false;
            double hi = 0.0;
            if (running) {
                if (nstep >= p.max_steps) {
                    rc = PV_ERR_MAX_STEPS;
                    running = false;
                } else {
                    ++nstep;
                    if (t + 1.01 * h >= tend) {
                        h = tend - t;
                        last = true;
                    }
                    if (!(fabs(h) > 1e-14 * fmax(fabs(t), 1.0))) {
                        rc = PV_ERR_STEP_TOO_SMALL;
                        running = false;
                    }
                    hi = pv_rcp_newton(h);
                }
            }
            // tick 0: state at the start of the step, Jacobian, LU of W = I/(gamma h) - J
            if (warp == 0 && running) {
                double J[M::NNZ], Wr[M::NLU];
                M::jac(t, z, C, J);
                M::lu_assemble(hi * (1.0 / GAMMA), J, Wr);
                M::lu_factor(Wr);
#pragma unroll
                for (int i = 0; i < M::NLU; ++i) Wsm.set(i, Wr[i]);
#pragma unroll
                for (int i = 0; i < NX; ++i) YN.set(i, z[i]);
            }
            __syncthreads();

            double zs[NX], r[NX];    // stage point of this block; right-hand side / stage vector being solved
            // entries of H_k (second-order term of the block-triangular Jacobian, at the start of the step): once per step
            double hc[M::NHC];
            if (warp > 0 && running) {
                if (warp == 1) M::template sens_term_prep<0>(t, YN, z, C, DC, hc);
                if constexpr (NS > 1)
                    if (warp == 2) M::template sens_term_prep<(NS > 1 ? 1 : 0)>(t, YN, z, C, DC, hc);
            }
            // Ticks 1..7: warp 0 does stage `tick`, the sensitivity warps stage `tick - 1`.  A real loop (the stage is a
            // run-time value): the role code - stage sums, RHS, H_k K^y, LU solve - exists once and stays in the
            // instruction cache.
#pragma unroll 1
            for (int tick = 1; tick <= 7; ++tick) {
                const int st = (warp == 0) ? tick : tick - 1;
                if (running && st >= 1 && st <= 6) {
                    const pv_smem_vec YS{sh + L::YS + ((st & 1) * NX) * 32 + lane};
                    // K_st of the state block: where warp 0 puts it and the sensitivity warps read it one tick later
                    double* const kyp = (st <= 5) ? sh + L::KS + ((st - 1) * NX) * 32 + lane : sh + L::K6Y + lane;
                    const pv_smem_vec KY{kyp};
                    const double ts = (st == 1) ? t : (st == 2) ? t + C2 * h : (st == 3) ? t + C3 * h : (st == 4) ? t + C4 * h : t + h;
                    // stage point zs = z + sum_j A[st][j] K_j and right-hand-side combination cs = sum_j C[st][j] K_j in one pass
                    // over this block's stage vectors (the newest K enters last)
                    // The first term initialises both sums (no copy of z, no zeroing); stage 1 has no K yet and sums over a
                    // row of zeros with zero coefficients instead: zs = z, cs = 0 exactly.
                    double cs[NX];
                    {
                        const double a = pv_rodas4_A[st - 1][0], cc = pv_rodas4_C[st - 1][0];
                        const double* k0 = (st == 1) ? sh + L::ZERO + lane : kmine;
#pragma unroll
                        for (int i = 0; i < NX; ++i) {
                            const double k = k0[i * 32];
                            zs[i] = fma(a, k, z[i]);
                            cs[i] = cc * k;
                        }
                    }
#pragma unroll 1
                    for (int j = 1; j < st - 1; ++j) {
                        const double a = pv_rodas4_A[st - 1][j], cc = pv_rodas4_C[st - 1][j];
                        const double* kj = kmine + (j * NX) * 32;
#pragma unroll
                        for (int i = 0; i < NX; ++i) {
                            const double k = kj[i * 32];
                            zs[i] = fma(a, k, zs[i]);
                            cs[i] = fma(cc, k, cs[i]);
                        }
                    }
                    if (warp == 0) {
#pragma unroll
                        for (int i = 0; i < NX; ++i) YS.set(i, zs[i]);
                        M::rhs(ts, zs, C, r);
                    } else {
                        if (warp == 1) M::template sens_rhs_block<0>(ts, YS, zs, C, DC, r);
                        if constexpr (NS > 1)
                            if (warp == 2) M::template sens_rhs_block<(NS > 1 ? 1 : 0)>(ts, YS, zs, C, DC, r);
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) r[i] = fma(hi, cs[i], r[i]);
                    if (warp > 0) {
                        if (warp == 1) M::template sens_term_apply<0>(hc, KY, C, DC, r);
                        if constexpr (NS > 1)
                            if (warp == 2) M::template sens_term_apply<(NS > 1 ? 1 : 0)>(hc, KY, C, DC, r);
                    }
                    M::lu_solve(Wsm, r);
                    if (st <= 5) {
                        double* kst = kmine + ((st - 1) * NX) * 32;   // for warp 0 this is kyp
#pragma unroll
                        for (int i = 0; i < NX; ++i) kst[i * 32] = r[i];
                    } else if (warp == 0) {
#pragma unroll
                        for (int i = 0; i < NX; ++i) KY.set(i, r[i]);
                    }
                    // K6 stays in r: new solution = stage point + K6, error estimate = K6
                }
                __syncthreads();
            }

            // ---- new solution zs + K6 (r holds K6), error partial of this block ----------------------------------
            double e2 = 0.0;
            bool finite = true;
            if (running) {
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const double znew = zs[i] + r[i];
                    const double sc = fma(p.rtol, pv_absmax(z[i], znew), p.atol);
                    const double q = r[i] * pv_rcp_approx(sc);
                    e2 = fma(q, q, e2);
                    finite = finite && (fabs(znew) < 1e300);
                    zs[i] = znew;
                }
                if (!finite || !(e2 == e2)) e2 = __longlong_as_double(0x7ff8000000000000LL);   // NaN marks a non-finite block
            }
            err_sh[warp * 32] = e2;
            __syncthreads();
            double err2d = 0.0;
#pragma unroll
            for (int w = 0; w < NB; ++w) err2d += err_sh[w * 32];
            float err2 = (float)(err2d * (1.0 / N));
            if (!(err2 <= 3.0e38f)) err2 = __int_as_float(0x7fc00000);
            bool accept = false;
            double tn = t;
            float fac = 1.f;
            if (running) {
                if (!(err2 == err2)) {
                    err2 = 1e10f;
                    if (fabs(h) < 1e-12) {
                        rc = PV_ERR_NONFINITE;
                        running = false;
                    }
                }
            }
            if (running) {
                const float err4 = pv_powf_fast(fmaxf(err2, 1e-30f), 0.125f);
                fac = fmaxf(1.0f / 6.0f, fminf(5.0f, err4 * (1.0f / 0.9f)));
                accept = err2 <= 1.0f;
                if (accept) {
                    ++n_acc;
                    tn = last ? tend : t + h;
                }
            }
            // ---- dense output (3rd order) for the grid points in (t, tn]; CTA-uniform loop ------------------------
            bool pending = accept && jout < T && sh_t[jout] <= tn;
            while (__syncthreads_or(pending)) {
                double zo[NX];
                if (pending) {
                    const double tj = sh_t[jout];
                    if (tj >= tn) {
#pragma unroll
                        for (int i = 0; i < NX; ++i) zo[i] = zs[i];
                    } else {
                        const double s = (tj - t) * hi, s1 = 1.0 - s;
#pragma unroll
                        for (int i = 0; i < NX; ++i) {
                            const double k1 = kmine[i * 32], k2 = kmine[(NX + i) * 32], k3 = kmine[(2 * NX + i) * 32],
                                         k4 = kmine[(3 * NX + i) * 32], k5 = kmine[(4 * NX + i) * 32];
                            const double c2 = fma(D21, k1, fma(D22, k2, fma(D23, k3, fma(D24, k4, D25 * k5))));
                            const double c3 = fma(D31, k1, fma(D32, k2, fma(D33, k3, fma(D34, k4, D35 * k5))));
                            zo[i] = fma(s1, z[i], s * fma(s1, fma(s, c3, c2), zs[i]));
                        }
                    }
                }
                sink(pending, zo);
                if (pending) {
                    ++jout;
                    pending = jout < T && sh_t[jout] <= tn;
                }
            }
            // ---- controller (Gustafsson, rodas.f PRED), replicated ----------------------------------------------
            if (running) {
                if (accept) {
                    if (n_acc > 1) {
                        float facgus = (float)(h_acc * hi) * pv_powf_fast(pv_fdiv_fast(fmaxf(err2, 1e-30f), err_acc), 0.25f) * (1.0f / 0.9f);
                        facgus = fmaxf(1.0f / 6.0f, fminf(5.0f, facgus));
                        fac = fmaxf(fac, facgus);
                    }
                    h_acc = h;
                    err_acc = fmaxf(1e-2f, sqrtf(err2));
                    double hnew = h * (double)pv_fdiv_fast(1.0f, fac);
                    if (last_rejected) hnew = fmin(hnew, h);
                    last_rejected = false;
#pragma unroll
                    for (int i = 0; i < NX; ++i) z[i] = zs[i];
                    t = tn;
                    if (last || jout >= T) {
                        rc = PV_OK;
                        running = false;
                    }
                    h = hnew;
                } else {
                    ++n_rej;
                    h = (n_acc == 0 && n_rej >= 2) ? h * 0.1 : h * (double)pv_fdiv_fast(1.0f, fac);
                    last_rejected = true;
                }
            }
        }

        // ---- results -------------------------------------------------------------------------------------------
        if (valid) {
            if (warp == 0) {
                p.logp[b] = (rc == 0) ? lp : __longlong_as_double(0x7ff8000000000000LL);
                if (p.status) p.status[b] = rc;
                if (p.steps) {
                    p.steps[b] = n_acc;
                    p.steps[p.B + b] = n_rej;
                }
            } else {
                // (the controller state, rc included, is replicated in every warp) a trajectory that did not finish has no
                // likelihood: NaN in logp, gradient and Fisher information alike
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                p.grad[(int64_t)(warp - 1) * p.B + b] = (rc == 0) ? g : nan;
                if constexpr (FIM) {
                    // upper triangle, row-major: row k = warp - 1 starts at k NS - k (k - 1) / 2
                    const int k = warp - 1, q0 = k * NS - k * (k - 1) / 2;
#pragma unroll
                    for (int r2 = 0; r2 < NS; ++r2)
                        if (r2 >= k) p.fim[(int64_t)(q0 + r2 - k) * p.B + b] = (rc == 0) ? fim[r2] : nan;
                }
            }
        }
    }
}
