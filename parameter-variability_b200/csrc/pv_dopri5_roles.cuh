// Block-per-lane DOPRI5 with forward sensitivities, for SMALL batches (BASELINE config 3 as it is literally stated:
// 8 chains x 1000 individuals = 8000 trajectories).  There the launch is one wave, and its duration is the LONGEST
// trajectory integrated by one lane at single-warp issue rate: 639 steps of a 9-component augmented system, ~2000 cycles
// each (0.65 - 0.8 ms; 7 % of the fp64 pipe).  Here a trajectory is spread over G = 1 + NS NEIGHBOURING LANES of a warp:
// lane 0 of the group integrates the state block, lane k the k-th sensitivity block - every lane a system of NX
// components instead of NX (1 + NS) - and the warp advances in lockstep (one step size, pv_dopri5_warp.cuh), so the stage
// state the sensitivity lanes need is one shuffle away and nothing diverges:
//   * all lanes run the SAME instruction stream: the generated role-uniform block function out = J(y) v + (df/dc)(y) w
//     (codegen emit.py: jv_plus_dfdc) with per-lane v and w; for models whose right-hand side is linear in the hoisted
//     coefficients the state lane uses it too (v = 0, w = c), otherwise it selects rhs(y);
//   * the error norm of a trajectory is the sum over its group's lanes (G shuffles), then the warp maximum (one REDUX);
//   * the likelihood folds into ONE accumulator per lane: the state lane's is logp, sensitivity lane k's is d logp / d theta_k.
// The dependent chain per step shrinks by ~(1 + NS); a warp carries 32 / G trajectories.  Same method, tableau,
// controller and dense output as the other DOPRI5 kernels; results agree with them within the integration tolerance.
// Replaces, like them, the AMICI forward-sensitivity solve behind pyPESTO's objective (petab_optimization.py:144-151).
#pragma once
#include "pv_dopri5_warp.cuh"

template <class M, bool QSH>
__global__ void __launch_bounds__(PV_BLOCK, 4)
pv_batch_roles_kernel(const __grid_constant__ pv_launch_params_m<M> p) {
    constexpr int NX = M::NX, NS = M::NS, G = 1 + NS, TPW = 32 / G;
    static_assert(NS >= 1 && G <= 32, "needs at least one sensitivity parameter");
    extern __shared__ __align__(16) double sh[];
    constexpr bool stats_shared = QSH;     // QSH <=> n_data == 1 (the host dispatches on it)
    double* sh_stats = sh;
    double* sh_t = sh + (stats_shared ? PV_QSTRIDE * p.T * p.n_obs : 0);
    double* sh_theta = sh_t + p.T;
    double* sh_y0 = sh_theta + M::NTH;
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) sh_t[i] = p.tgrid[i];
    for (int i = threadIdx.x; i < M::NTH; i += blockDim.x) sh_theta[i] = p.theta_base[i];
    for (int i = threadIdx.x; i < NX; i += blockDim.x) sh_y0[i] = p.y0_base[i];
    if (stats_shared)
        for (int i = threadIdx.x; i < PV_QSTRIDE * p.T * p.n_obs; i += blockDim.x) sh_stats[i] = p.stats[i];
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    // lanes beyond the last whole group repeat the last group's state lane (identical values to identical addresses)
    const int tr = min(lane_id / G, TPW - 1);
    const int role = lane_id < TPW * G ? lane_id - tr * G : 0;
    const int src = tr * G;
    const int64_t n_items = p.n_items_dev ? (int64_t)*p.n_items_dev : p.B;
    const int T = p.T;
    const double rtol = p.rtol, atol = p.atol;
    const double tend = sh_t[T - 1];

    for (;;) {
        unsigned long long base = 0;
        if (lane_id == 0) base = atomicAdd(p.counter, (unsigned long long)TPW);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= n_items) break;
        const int64_t item = min((int64_t)base + tr, n_items - 1);
        const int64_t b = p.index_map ? (int64_t)p.index_map[item] : item;

        // ---- set-up (every lane of a group computes its trajectory's coefficients) ------------------------------
        double c[M::NC], w[M::NC], obs[NX];
        double v[2][NX], k1[2][NX];
        {
            double pc[PV_MAX_COLS];
#pragma unroll
            for (int j = 0; j < PV_MAX_COLS; ++j) pc[j] = (j < p.n_cols) ? __ldg(p.P + (int64_t)j * p.B + b) : 0.0;
            double th[M::NTH];
#pragma unroll
            for (int i = 0; i < M::NTH; ++i) {
                double x = sh_theta[i];
#pragma unroll
                for (int j = 0; j < PV_MAX_COLS; ++j) x = (j < p.n_cols && p.col_target[j] == i) ? pc[j] : x;
                th[i] = x;
            }
            double y0[NX], dc[M::NDC], dy0[M::NSD][NX];
            M::hoist(th, c, y0, obs);
            M::hoist_sens(th, dc, dy0);
            M::dc_dense(role - 1, dc, w);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double x = sh_y0[i];
                bool over = (x == x);
                x = over ? x : y0[i];
#pragma unroll
                for (int j = 0; j < PV_MAX_COLS; ++j)
                    if (j < p.n_cols && p.col_target[j] == M::NTH + i) {
                        x = pc[j];
                        over = true;
                    }
                double s0 = 0.0;
#pragma unroll
                for (int s = 0; s < NS; ++s) s0 = (role == s + 1 && !over) ? dy0[s][i] : s0;
                v[0][i] = role == 0 ? x : s0;
            }
            if constexpr (M::RHS_LINEAR_IN_C) {     // the state lane evaluates f(y) = (df/dc)(y) c through the same function
#pragma unroll
                for (int m = 0; m < M::NC; ++m) w[m] = role == 0 ? c[m] : w[m];
            }
        }
        // block right-hand side at stage state: the group's state lane holds y
        auto block_rhs = [&](double ts, const double (&vs)[NX], double (&out)[NX]) {
            double ys[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) ys[i] = __shfl_sync(FULL, vs[i], src);
            if constexpr (M::RHS_LINEAR_IN_C) {
                double vv[NX];
#pragma unroll
                for (int i = 0; i < NX; ++i) vv[i] = role == 0 ? 0.0 : vs[i];
                M::jv_plus_dfdc(ts, ys, vv, c, w, out);
            } else {
                double fy[NX], jo[NX];
                M::rhs(ts, ys, c, fy);
                M::jv_plus_dfdc(ts, ys, vs, c, w, jo);
#pragma unroll
                for (int i = 0; i < NX; ++i) out[i] = role == 0 ? fy[i] : jo[i];
            }
        };
        // sum over the lanes of a group
        auto group_sum = [&](double x) {
            double tot = 0.0;
#pragma unroll
            for (int r = 0; r < G; ++r) tot += __shfl_sync(FULL, x, src + r);
            return tot;
        };

        // ---- likelihood accumulator: logp in the state lane, d logp / d theta_k in sensitivity lane k ----------------
        double acc = 0.0;
        const double* qrow = nullptr;
        unsigned qaddr = 0;
        if constexpr (QSH)
            qaddr = (unsigned)__cvta_generic_to_shared(sh_stats);
        else
            qrow = p.stats + (int64_t)(b % p.n_data) * (T * p.n_obs * PV_QSTRIDE);
        auto output = [&](const double (&vo)[NX]) {
            double fo[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) fo[i] = __shfl_sync(FULL, vo[i], src);
            for (int k = 0; k < p.n_obs; ++k) {
                const int idx = p.obs_idx[k];
                double sc = 0.0, f = 0.0, df = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    sc = idx == i ? obs[i] : sc;
                    f = idx == i ? fo[i] : f;
                    df = idx == i ? vo[i] : df;
                }
                f *= sc;
                df *= sc;
                double q0, q1, q2;
                if constexpr (QSH) {
                    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q0), "=d"(q1) : "r"(qaddr + k * (PV_QSTRIDE * 8)));
                    asm("ld.shared.f64 %0, [%1];" : "=d"(q2) : "r"(qaddr + k * (PV_QSTRIDE * 8) + 16));
                } else {
                    const double2 q01 = *reinterpret_cast<const double2*>(qrow + k * PV_QSTRIDE);
                    q2 = qrow[k * PV_QSTRIDE + 2];
                    q0 = q01.x;
                    q1 = q01.y;
                }
                if (p.noise_model == 0) {
                    acc += role == 0 ? fma(f, fma(q2, f, q1), q0) : fma(2.0 * q2, f, q1) * df;
                } else if (q2 != 0.0) {
                    const double lf = log(f);
                    acc += role == 0 ? fma(lf, fma(q2, lf, q1), q0) : fma(2.0 * q2, lf, q1) * df / f;
                }
            }
            if constexpr (QSH)
                qaddr += p.n_obs * (PV_QSTRIDE * 8);
            else
                qrow += p.n_obs * PV_QSTRIDE;
        };

        // ---- start: outputs at t0, first derivative, initial step (the smallest of the warp) -------------------------
        double t = p.t0, h;
        int jout = 0, nstep = 0, n_acc = 0, n_rej = 0, cur = 0, rc = PV_RUNNING, stiff_run = 0, nonfinite_run = 0;
        bool dead = false, last_rejected = false;
        float lfacold = -13.287712f;
        while (jout < T && sh_t[jout] <= t) {
            output(v[0]);
            ++jout;
        }
        double t_next = jout < T ? sh_t[jout] : 1e300, t_next1 = jout + 1 < T ? sh_t[jout + 1] : 1e300;
        if (jout >= T) rc = PV_OK;
        block_rhs(t, v[0], k1[0]);
        {
            double d0 = 0.0, d1 = 0.0, yn[NX], k2[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double sc = atol + rtol * fabs(v[0][i]);
                const double a = v[0][i] / sc, bb = k1[0][i] / sc;
                d0 += a * a;
                d1 += bb * bb;
            }
            d0 = sqrt(group_sum(d0) / (NX * G));
            d1 = sqrt(group_sum(d1) / (NX * G));
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            h0 = fmin(fmax(h0, 1e-12), tend - t);
#pragma unroll
            for (int i = 0; i < NX; ++i) yn[i] = fma(h0, k1[0][i], v[0][i]);
            block_rhs(t + h0, yn, k2);
            double d2 = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double sc = atol + rtol * fabs(v[0][i]);
                const double a = (k2[i] - k1[0][i]) / sc;
                d2 += a * a;
            }
            d2 = sqrt(group_sum(d2) / (NX * G)) / h0;
            const double dm = fmax(d1, d2);
            const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
            h = fmin(fmin(100.0 * h0, h1), tend - t);
            const bool sane = (d0 == d0) && (d1 == d1) && (d2 == d2) && (h == h);
            h = fmin(pv_warp_min(sane ? h : 1e300), tend - t);
        }

        // ---- steps: register set `cur` holds (v, k1) at t; an accepted step flips it -----------------------------------
        auto step = [&](auto cur_tag) {
            constexpr int CUR = decltype(cur_tag)::value;
            double(&Y)[NX] = v[CUR];
            double(&K1)[NX] = k1[CUR];
            double(&YN)[NX] = v[1 - CUR];
            double(&K7)[NX] = k1[1 - CUR];
            if (nstep >= p.max_steps) return (int)PV_ERR_MAX_STEPS;
            ++nstep;
            bool last = false;
            if (t + 1.01 * h >= tend) {
                h = tend - t;
                last = true;
            }
            double k2[NX], k3[NX], k4[NX], k5[NX], k6[NX], yn[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) yn[i] = fma(h * pv_dp5[0], K1[i], Y[i]);
            block_rhs(t + 0.2 * h, yn, k2);
#pragma unroll
            for (int i = 0; i < NX; ++i) yn[i] = fma(h, fma(pv_dp5[2], k2[i], pv_dp5[1] * K1[i]), Y[i]);
            block_rhs(t + 0.3 * h, yn, k3);
#pragma unroll
            for (int i = 0; i < NX; ++i) yn[i] = fma(h, fma(pv_dp5[5], k3[i], fma(pv_dp5[4], k2[i], pv_dp5[3] * K1[i])), Y[i]);
            block_rhs(t + 0.8 * h, yn, k4);
#pragma unroll
            for (int i = 0; i < NX; ++i)
                yn[i] = fma(h, fma(pv_dp5[9], k4[i], fma(pv_dp5[8], k3[i], fma(pv_dp5[7], k2[i], pv_dp5[6] * K1[i]))), Y[i]);
            block_rhs(t + (8.0 / 9.0) * h, yn, k5);
#pragma unroll
            for (int i = 0; i < NX; ++i)
                yn[i] = fma(h, fma(pv_dp5[14], k5[i], fma(pv_dp5[13], k4[i], fma(pv_dp5[12], k3[i], fma(pv_dp5[11], k2[i], pv_dp5[10] * K1[i])))),
                            Y[i]);
            block_rhs(t + h, yn, k6);
#pragma unroll
            for (int i = 0; i < NX; ++i)
                YN[i] = fma(h, fma(pv_dp5[19], k6[i], fma(pv_dp5[18], k5[i], fma(pv_dp5[17], k4[i], fma(pv_dp5[16], k3[i], pv_dp5[15] * K1[i])))),
                            Y[i]);
            block_rhs(t + h, YN, K7);
            // error: this lane's block, summed over the group, maximum over the warp
            double e2 = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double e = h * fma(pv_dp5[25], K7[i],
                                         fma(pv_dp5[24], k6[i], fma(pv_dp5[23], k5[i], fma(pv_dp5[22], k4[i], fma(pv_dp5[21], k3[i], pv_dp5[20] * K1[i])))));
                const double sc = atol + rtol * fmax(fabs(Y[i]), fabs(YN[i]));
                const double r = e * pv_rcp_approx(sc);
                e2 = fma(r, r, e2);
            }
            float err2 = (float)(group_sum(e2) * (1.0 / (NX * G)));
            const bool bad = !(err2 <= 3.0e38f);
            nonfinite_run = bad ? nonfinite_run + 1 : 0;
            dead = dead || nonfinite_run >= 8;
            err2 = pv_warp_max_nonneg(dead ? 0.0f : (bad ? 1e10f : err2));
            if (err2 >= 1e10f && fabs(h) < 1e-12) return (int)PV_ERR_NONFINITE;
            const float l2 = pv_log2f_fast(fmaxf(err2, 1e-30f));
            const float lfac11 = 0.085f * l2;
            if (err2 <= 1.0f) {
                ++n_acc;
                const double tn = last ? tend : t + h;
                if (t_next <= tn) {
                    double yd[NX], bs[NX], r4[NX], r5[NX];
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        yd[i] = YN[i] - Y[i];
                        bs[i] = fma(h, K1[i], -yd[i]);
                        r4[i] = yd[i] - fma(h, K7[i], bs[i]);
                        r5[i] = h * fma(pv_dp5[31], K7[i],
                                        fma(pv_dp5[30], k6[i], fma(pv_dp5[29], k5[i], fma(pv_dp5[28], k4[i], fma(pv_dp5[27], k3[i], pv_dp5[26] * K1[i])))));
                    }
                    const double hinv = pv_rcp_newton(h);
                    do {
                        const double th = (t_next >= tn) ? 1.0 : (t_next - t) * hinv, th1 = 1.0 - th;
                        double vo[NX];
#pragma unroll
                        for (int i = 0; i < NX; ++i) vo[i] = fma(th, fma(th1, fma(th, fma(th1, r5[i], r4[i]), bs[i]), yd[i]), Y[i]);
                        output(vo);
                        ++jout;
                        t_next = t_next1;
                        t_next1 = jout + 1 < T ? sh_t[jout + 1] : 1e300;
                    } while (t_next <= tn);
                }
                if (nstep > p.stiff_after && (nstep & 3) == 0) {
                    double num = 0.0, den = 0.0;
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const double g6 = fma(h, fma(pv_dp5[14], k5[i], fma(pv_dp5[13], k4[i], fma(pv_dp5[12], k3[i], fma(pv_dp5[11], k2[i], pv_dp5[10] * K1[i])))), Y[i]);
                        const double dy = YN[i] - g6, dk = K7[i] - k6[i];
                        num = fma(dk, dk, num);
                        den = fma(dy, dy, den);
                    }
                    if (pv_warp_any(!dead && den > 0.0 && h * h * num > 10.5625 * den)) {
                        stiff_run = (stiff_run & ~0xffff) + 0x10000;
                        if (stiff_run >= 15 * 0x10000 && (tend - t) > PV_STIFF_REMAINING_STEPS * h) return (int)PV_ERR_STIFF;
                    } else if (((++stiff_run) & 0xffff) >= 6) {
                        stiff_run = 0;
                    }
                }
                const float lfac = fminf(fmaxf(lfac11 - 0.04f * lfacold + 0.15200309f, last_rejected ? 0.0f : -3.3219281f), 2.3219281f);
                const double hnew = h * (double)pv_exp2f_fast(-lfac);
                lfacold = fmaxf(0.5f * l2, -13.287712f);
                last_rejected = false;
                cur = 1 - CUR;
                t = tn;
                if (last || jout >= T) return (int)PV_OK;
                h = hnew;
            } else {
                ++n_rej;
                h = h * (double)pv_exp2f_fast(-fminf(lfac11 + 0.15200309f, 2.3219281f));
                last_rejected = true;
                if (!(fabs(h) > 1e-14 * fmax(fabs(t), 1.0))) return (int)PV_ERR_STEP_TOO_SMALL;
            }
            return (int)PV_RUNNING;
        };
        while (rc == PV_RUNNING) rc = cur ? step(pv_int_tag<1>{}) : step(pv_int_tag<0>{});
        // a trajectory is dead if any lane of its group is
        const double n_dead = group_sum(dead ? 1.0 : 0.0);      // (every lane takes part in the shuffles: no short-circuit)
        if (n_dead > 0.0) rc = PV_ERR_NONFINITE;

        // ---- results ---------------------------------------------------------------------------------------------------
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (role == 0) {
            p.logp[b] = (rc == 0) ? acc : nan;
            if (p.status) p.status[b] = rc;
            if (p.steps) {
                p.steps[b] = n_acc;
                p.steps[p.B + b] = n_rej;
            }
        } else {
            p.grad[(int64_t)(role - 1) * p.B + b] = (rc == 0) ? acc : nan;
        }
    }
}
