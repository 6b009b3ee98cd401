// Dormand-Prince 5(4) for a WARP in lockstep (north_star subsystem 2: "warp-level vote / shuffle for step-size agreement
// and divergence control").  The 32 lanes integrate 32 trajectories - neighbours in parameter space under the locality
// schedule (pv_schedule.cuh) - with ONE step size: the controller sees the largest error norm of the warp (one REDUX), so
// t, h, accept / reject, the grid crossings, start and end are warp-uniform.  Consequences the per-lane stepper
// (pv_dopri5.cuh) cannot have:
//   * no divergence and no refill: the per-grid-point code runs with all 32 lanes or not at all;
//   * (y, k1) and (y_new, k7) PING-PONG between two register sets chosen by a compile-time index: an accepted step flips
//     the index (a warp-uniform branch to the other copy of the step) instead of moving 2 N doubles - the per-lane
//     stepper spent 16 % of its instructions on those moves;
//   * the next grid time is prefetched one output ahead, and the likelihood coefficients are read from shared memory with
//     ld.shared (no aliasing with the trajectory stores, so the loads overlap the interpolant).
// Every lane still meets its own tolerance (its error <= the warp's).  Same method, tableau, controller, dense output and
// stiffness test as pv_dopri5_lane; replaces the reference's per-row CVODE call (petab_factory.py:242-246) likewise.
#pragma once
#include <type_traits>
#include "pv_dopri5.cuh"

// A system may spread ONE trajectory over Sys::GROUP neighbouring lanes (pv_dopri5_pairs.cuh): every lane of the group
// carries the first Sys::ERR_SHARED components (the state block), the rest are the lane's own; the trajectory's error norm is
// the mean square over its Sys::ERR_N distinct components.
template <class Sys, class = void>
struct pv_group_of {
    static constexpr int G = 0;      // not grouped: one trajectory per lane
    static constexpr int ERR_N = Sys::N;
};
template <class Sys>
struct pv_group_of<Sys, std::void_t<decltype(Sys::GROUP)>> {
    static constexpr int G = Sys::GROUP;
    static constexpr int ERR_N = Sys::ERR_N;
};

template <class Sys>
struct pv_dopri5_warp {
    static constexpr int N = Sys::N;
    static constexpr int G = pv_group_of<Sys>::G;
    int group_src = 0;           // first lane of this lane's group (grouped systems)
    // sum over the lanes of this lane's group
    PV_D double group_sum(double x) const {
#if defined(__CUDA_ARCH__)
        if constexpr (G > 1) {
            double tot = 0.0;
#pragma unroll
            for (int r = 0; r < G; ++r) tot += __shfl_sync(0xffffffffu, x, group_src + r);
            return tot;
        }
#endif
        return x;
    }
    double t, h, tend;
    double t_next, t_next1;      // tgrid[jout], tgrid[jout + 1] (1e300 past the end): the second is loaded one output ahead
    double y[2][N], k1[2][N];    // [cur] = state and its derivative at t; [1 - cur] receives y_new / k7 of the step
    float lfacold;
    bool last_rejected;
    int cur, jout, nstep;
    pv_step_stats st;
    int stiff_after = 0x7fffffff;
    int stiff_run;
    // a lane whose own error estimate stays non-finite leaves the agreement instead of dragging its neighbours' step to 0
    int nonfinite_run;
    bool dead;

    template <class Out>
    PV_D int start(const typename Sys::Consts& ks, const double (&z0)[N], double t0, const double* tgrid, int T, double rtol,
                   double atol, Out& out) {
        st.accepted = st.rejected = 0;
        nstep = 0;
        stiff_run = 0;
        nonfinite_run = 0;
        dead = false;
        cur = 0;
        lfacold = -13.287712f;   // log2(1e-4)
        last_rejected = false;
        t = t0;
        tend = tgrid[T - 1];
        h = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) y[0][i] = z0[i];
        jout = 0;
        while (jout < T && tgrid[jout] <= t) {   // uniform: every lane has the same t0 and grid
            out(jout, tgrid[jout], y[0]);
            ++jout;
        }
        if (jout >= T) return PV_OK;
        t_next = tgrid[jout];
        t_next1 = jout + 1 < T ? tgrid[jout + 1] : 1e300;
        Sys::rhs(t, y[0], ks, k1[0]);
        // initial step (Hairer, Norsett, Wanner I, II.4) per lane, then the smallest of the warp
        double yn[N], k2[N];
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sc = atol + rtol * fabs(y[0][i]);
            const double a = y[0][i] / sc, b = k1[0][i] / sc;
            d0 += a * a;
            d1 += b * b;
        }
        d0 = sqrt(d0 / N);
        d1 = sqrt(d1 / N);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        h0 = fmin(fmax(h0, 1e-12), tend - t);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h0, k1[0][i], y[0][i]);
        Sys::rhs(t + h0, yn, ks, k2);
        double d2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sc = atol + rtol * fabs(y[0][i]);
            const double a = (k2[i] - k1[0][i]) / sc;
            d2 += a * a;
        }
        d2 = sqrt(d2 / N) / h0;
        const double dm = fmax(d1, d2);
        const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
        h = fmin(fmin(100.0 * h0, h1), tend - t);
        // a lane whose estimate is not a number (NaN parameters) does not vote: it will be `dead` within 8 steps, and fmin /
        // fmax would have turned its NaNs into the lower clamp 1e-12
        const bool sane = (d0 == d0) && (d1 == d1) && (d2 == d2) && (h == h);
        h = pv_warp_min(sane ? h : 1e300);
        h = fmin(h, tend - t);
        return PV_RUNNING;
    }

    // one attempted step of the whole warp, reading register set CUR and writing set 1 - CUR
    template <int CUR, class Out>
    PV_D int step(const typename Sys::Consts& ks, const double* tgrid, int T, double rtol, double atol, int max_steps, Out& out) {
        double(&Y)[N] = y[CUR];
        double(&K1)[N] = k1[CUR];
        double(&YN)[N] = y[1 - CUR];
        double(&K7)[N] = k1[1 - CUR];
        if (nstep >= max_steps) return PV_ERR_MAX_STEPS;
        ++nstep;
        bool last = false;
        if (t + 1.01 * h >= tend) {
            h = tend - t;
            last = true;
        }
        double k2[N], k3[N], k4[N], k5[N], k6[N], yn[N];
        // stages: the newest derivative enters each sum last (see pv_dopri5_lane::step)
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h * pv_dp5[0], K1[i], Y[i]);
        Sys::rhs(t + 0.2 * h, yn, ks, k2);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h, fma(pv_dp5[2], k2[i], pv_dp5[1] * K1[i]), Y[i]);
        Sys::rhs(t + 0.3 * h, yn, ks, k3);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h, fma(pv_dp5[5], k3[i], fma(pv_dp5[4], k2[i], pv_dp5[3] * K1[i])), Y[i]);
        Sys::rhs(t + 0.8 * h, yn, ks, k4);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = fma(h, fma(pv_dp5[9], k4[i], fma(pv_dp5[8], k3[i], fma(pv_dp5[7], k2[i], pv_dp5[6] * K1[i]))), Y[i]);
        Sys::rhs(t + (8.0 / 9.0) * h, yn, ks, k5);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = fma(h, fma(pv_dp5[14], k5[i], fma(pv_dp5[13], k4[i], fma(pv_dp5[12], k3[i], fma(pv_dp5[11], k2[i], pv_dp5[10] * K1[i])))),
                        Y[i]);
        Sys::rhs(t + h, yn, ks, k6);
#pragma unroll
        for (int i = 0; i < N; ++i)
            YN[i] = fma(h, fma(pv_dp5[19], k6[i], fma(pv_dp5[18], k5[i], fma(pv_dp5[17], k4[i], fma(pv_dp5[16], k3[i], pv_dp5[15] * K1[i])))),
                        Y[i]);
        Sys::rhs(t + h, YN, ks, K7);

        // error estimate: fp32 norm of an fp64 difference, then the warp's maximum
        double err2d = 0.0, err2own = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double e = h * fma(pv_dp5[25], K7[i],
                                     fma(pv_dp5[24], k6[i], fma(pv_dp5[23], k5[i], fma(pv_dp5[22], k4[i], fma(pv_dp5[21], k3[i], pv_dp5[20] * K1[i])))));
            const double sc = atol + rtol * pv_absmax(Y[i], YN[i]);
            const double r = e * pv_rcp_approx(sc);
            if constexpr (G > 0) {
                if (i < Sys::ERR_SHARED)
                    err2d = fma(r, r, err2d);
                else
                    err2own = fma(r, r, err2own);
            } else {
                err2d = fma(r, r, err2d);
            }
        }
        float err2;
        if constexpr (G > 0)
            err2 = (float)((err2d + group_sum(err2own)) * (1.0 / pv_group_of<Sys>::ERR_N));
        else
            err2 = (float)(err2d * (1.0 / N));
        const bool bad = !(err2 <= 3.0e38f);
        nonfinite_run = bad ? nonfinite_run + 1 : 0;
        dead = dead || nonfinite_run >= 8;
        err2 = pv_warp_max_nonneg(dead ? 0.0f : (bad ? 1e10f : err2));
        if (err2 >= 1e10f && fabs(h) < 1e-12) return PV_ERR_NONFINITE;
        const float l2 = pv_log2f_fast(fmaxf(err2, 1e-30f));
        const float lfac11 = 0.085f * l2;
        if (err2 <= 1.0f) {
            ++st.accepted;
            const double tn = last ? tend : t + h;
            if (t_next <= tn) {
                // y(t + th h) = y + th (yd + (1 - th) (bs + th (r4 + (1 - th) r5)))   (Hairer's contd5, Horner form)
                double yd[N], bs[N], r4[N], r5[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    yd[i] = YN[i] - Y[i];
                    bs[i] = fma(h, K1[i], -yd[i]);
                    r4[i] = yd[i] - fma(h, K7[i], bs[i]);
                    r5[i] = h * fma(pv_dp5[31], K7[i],
                                    fma(pv_dp5[30], k6[i], fma(pv_dp5[29], k5[i], fma(pv_dp5[28], k4[i], fma(pv_dp5[27], k3[i], pv_dp5[26] * K1[i])))));
                }
                const double hinv = pv_rcp_newton(h);
                do {
                    const double th = (t_next >= tn) ? 1.0 : (t_next - t) * hinv, th1 = 1.0 - th;
                    double yo[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) yo[i] = fma(th, fma(th1, fma(th, fma(th1, r5[i], r4[i]), bs[i]), yd[i]), Y[i]);
                    out(jout, t_next, yo);
                    ++jout;
                    t_next = t_next1;
                    t_next1 = jout + 1 < T ? tgrid[jout + 1] : 1e300;      // needed one output from now
                } while (t_next <= tn);
            }
            if (nstep > stiff_after && (nstep & 3) == 0) {      // every 4th step: the test costs ~30 flops + a vote
                double num = 0.0, den = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double g6 = fma(h, fma(pv_dp5[14], k5[i], fma(pv_dp5[13], k4[i], fma(pv_dp5[12], k3[i], fma(pv_dp5[11], k2[i], pv_dp5[10] * K1[i])))), Y[i]);
                    const double dy = YN[i] - g6, dk = K7[i] - k6[i];
                    num = fma(dk, dk, num);
                    den = fma(dy, dy, den);
                }
                if (pv_warp_any(!dead && den > 0.0 && h * h * num > 10.5625 * den)) {
                    stiff_run = (stiff_run & ~0xffff) + 0x10000;
                    if (stiff_run >= 15 * 0x10000 && (tend - t) > PV_STIFF_REMAINING_STEPS * h) return PV_ERR_STIFF;
                } else if (((++stiff_run) & 0xffff) >= 6) {
                    stiff_run = 0;
                }
            }
            const float lfac = fminf(fmaxf(lfac11 - 0.04f * lfacold + 0.15200309f, last_rejected ? 0.0f : -3.3219281f), 2.3219281f);
            const double hnew = h * (double)pv_exp2f_fast(-lfac);
            lfacold = fmaxf(0.5f * l2, -13.287712f);
            last_rejected = false;
            cur = 1 - CUR;                       // the roles swap: no copy of y_new / k7
            t = tn;
            if (last || jout >= T) return PV_OK;
            h = hnew;
        } else {
            ++st.rejected;
            h = h * (double)pv_exp2f_fast(-fminf(lfac11 + 0.15200309f, 2.3219281f));
            last_rejected = true;
            if (!(fabs(h) > 1e-14 * fmax(fabs(t), 1.0))) return PV_ERR_STEP_TOO_SMALL;
        }
        return PV_RUNNING;
    }
};
