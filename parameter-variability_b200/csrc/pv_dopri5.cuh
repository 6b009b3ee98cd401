// Adaptive Dormand-Prince 5(4) with dense output, one trajectory per thread (north_star subsystem 2,
// non-stiff path).  Replaces, for many parameter rows at once, the per-row
// `RoadRunner.simulate(start, end, steps)` CVODE call of the reference
// (src/parvar/experiments/petab_factory.py:242-246) and, with the augmented sensitivity system,
// the CVODES forward-sensitivity solve behind pyPESTO's objective
// (src/parvar/optimization/petab_optimization.py:144-151).
//
// Everything a lane needs lives in registers: state, 7 stage derivatives, hoisted model
// coefficients.  The step-size controller (error norm, PI factor) runs in fp32 on the SFU
// (`__powf`) - it only steers h; every quantity that enters the solution is fp64.
#pragma once
#include "pv_common.cuh"

// ---- system adaptors: what the integrators see ---------------------------------------------
template <class M, bool SENS>
struct pv_system;

template <class M>
struct pv_system<M, false> {
    static constexpr int NX = M::NX;
    static constexpr int N = M::NX;
    static constexpr int FLOPS_RHS = M::FLOPS_RHS;
    struct Consts {
        double c[M::NC];
    };
    PV_HD static void rhs(double t, const double (&z)[N], const Consts& k, double (&fz)[N]) {
        M::rhs(t, z, k.c, fz);
    }
};

template <class M>
struct pv_system<M, true> {
    static constexpr int NX = M::NX;
    static constexpr int N = M::NX * (1 + M::NS);
    static constexpr int FLOPS_RHS = M::FLOPS_RHS_SENS;
    struct Consts {
        double c[M::NC];
        double dc[M::NDC];
    };
    PV_HD static void rhs(double t, const double (&z)[N], const Consts& k, double (&fz)[N]) {
        double y[M::NX], s[M::NSD][M::NX], f[M::NX], fs[M::NSD][M::NX];
#pragma unroll
        for (int i = 0; i < M::NX; ++i) y[i] = z[i];
#pragma unroll
        for (int p = 0; p < M::NS; ++p)
#pragma unroll
            for (int i = 0; i < M::NX; ++i) s[p][i] = z[M::NX * (1 + p) + i];
        M::rhs_sens(t, y, s, k.c, k.dc, f, fs);
#pragma unroll
        for (int i = 0; i < M::NX; ++i) fz[i] = f[i];
#pragma unroll
        for (int p = 0; p < M::NS; ++p)
#pragma unroll
            for (int i = 0; i < M::NX; ++i) fz[M::NX * (1 + p) + i] = fs[p][i];
    }
};

// ---- controller helpers -------------------------------------------------------------------
PV_HD float pv_log2f_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;                     // bare MUFU.LG2: callers clamp the argument away from the denormal range
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return log2f(x);
#endif
}
PV_HD float pv_exp2f_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;                     // bare MUFU.EX2: the argument here is clamped to [-3.4, 2.4], no range scaling needed
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f(x);
#endif
}
PV_HD float pv_powf_fast(float x, float a) {
#if defined(__CUDA_ARCH__)
    return __powf(x, a);
#else
    return powf(x, a);
#endif
}

PV_HD double pv_drcp(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

// Butcher tableau of DOPRI5 in __constant__ memory: fp64 literals cannot be instruction immediates, so the
// compiler would otherwise rebuild every coefficient with two UMOVs per use; a c[bank][offset] operand is free.
#if defined(__CUDACC__)
#define PV_COEF_SPACE __constant__
#else
#define PV_COEF_SPACE static const
#endif
PV_COEF_SPACE double pv_dp5[32] = {
    /* 0 a21 */ 1.0 / 5.0,
    /* 1 a31 */ 3.0 / 40.0, /* 2 a32 */ 9.0 / 40.0,
    /* 3 a41 */ 44.0 / 45.0, /* 4 a42 */ -56.0 / 15.0, /* 5 a43 */ 32.0 / 9.0,
    /* 6 a51 */ 19372.0 / 6561.0, /* 7 a52 */ -25360.0 / 2187.0, /* 8 a53 */ 64448.0 / 6561.0, /* 9 a54 */ -212.0 / 729.0,
    /* 10 a61 */ 9017.0 / 3168.0, /* 11 a62 */ -355.0 / 33.0, /* 12 a63 */ 46732.0 / 5247.0, /* 13 a64 */ 49.0 / 176.0,
    /* 14 a65 */ -5103.0 / 18656.0,
    /* 15 b1 */ 35.0 / 384.0, /* 16 b3 */ 500.0 / 1113.0, /* 17 b4 */ 125.0 / 192.0, /* 18 b5 */ -2187.0 / 6784.0,
    /* 19 b6 */ 11.0 / 84.0,
    /* 20 e1 */ 71.0 / 57600.0, /* 21 e3 */ -71.0 / 16695.0, /* 22 e4 */ 71.0 / 1920.0, /* 23 e5 */ -17253.0 / 339200.0,
    /* 24 e6 */ 22.0 / 525.0, /* 25 e7 */ -1.0 / 40.0,
    /* 26 d1 */ -12715105075.0 / 11282082432.0, /* 27 d3 */ 87487479700.0 / 32700410799.0,
    /* 28 d4 */ -10690763975.0 / 1880347072.0, /* 29 d5 */ 701980252875.0 / 199316789632.0,
    /* 30 d6 */ -1453857185.0 / 822651844.0, /* 31 d7 */ 69997945.0 / 29380423.0};

#ifndef PV_STIFF_REMAINING_STEPS
#define PV_STIFF_REMAINING_STEPS 2500.0
#endif

// ---- warp agreement (lockstep schedule) ---------------------------------------------------------------------
// max over the warp of a non-negative float (one REDUX on the bit pattern: for x >= 0 the IEEE order is the integer order)
PV_HD float pv_warp_max_nonneg(float x) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(x)));
#else
    return x;
#endif
}
PV_HD double pv_warp_min(double x) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, off));
#endif
    return x;
}
PV_HD bool pv_warp_any(bool x) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, x);
#else
    return x;
#endif
}

struct pv_step_stats {
    int accepted;
    int rejected;
};

enum pv_status_code : int {
    PV_OK = 0,
    PV_ERR_MAX_STEPS = 1,    // step budget exhausted before t_end
    PV_ERR_STEP_TOO_SMALL = 2,
    PV_ERR_NONFINITE = 3,
    PV_ERR_STIFF = 4,        // explicit integrator gave up: step size limited by stability (see pv_dopri5_lane::stiff_after)
};

// Flops of one attempted step besides the RHS evaluations: 21 a_ij + 6 b/e FMAs for the stages
// and the error vector (2 flop each) plus ~10 per component for the norm, per state.
template <class Sys>
struct pv_dopri5_cost {
    static constexpr int STAGES = 6;   // FSAL: k7 of an accepted step is k1 of the next
    static constexpr int FLOPS_STEP = Sys::N * (2 * (1 + 2 + 3 + 4 + 5 + 5 + 6) + 10) + STAGES * Sys::FLOPS_RHS + 12;
    static constexpr int FLOPS_DENSE = Sys::N * (2 * 6 + 6 + 8);   // rcont5 + interpolant per output row
};

// status of a lane that still has work to do
#define PV_RUNNING (-1)

// Per-lane DOPRI5 state machine: start() once, then step() until it returns something other than
// PV_RUNNING.  The batch kernels drive it either one trajectory per thread for the whole launch
// (static schedule) or with lane refill (a lane that finishes picks the next trajectory while its
// warp neighbours keep stepping).  `out(j, t_j, z_j)` is called once per grid point in increasing j.
template <class Sys>
struct pv_dopri5_lane {
    static constexpr int N = Sys::N;
    double t, h, tend;
    double t_next;     // tgrid[jout]: next grid point not yet covered (kept in a register: checked every step)
    double y[N], k1[N];
    float lfacold;     // log2 of Hairer's facold = max(err, 1e-4)
    bool last_rejected;
    int jout, nstep;
    pv_step_stats st;
    // Stiffness detection (Hairer, Norsett, Wanner II.4 / dopri5.f): after `stiff_after` attempted steps every 4th accepted
    // step estimates h * |lambda| ~ h * |f(t+h, y1) - f(t+h, g6)| / |y1 - g6| from the last two stages; 15 consecutive
    // estimates above DOPRI5's stability bound 3.25 mean the step size is set by stability, not accuracy.  That alone
    // is not a reason to stop: simple_pk's tail (t > 40 of [0, 100], solution at atol level) is stability limited at
    // h ~ 1.3 and still only ~60 steps long.  The lane gives up (PV_ERR_STIFF, so the caller can hand the trajectory to
    // the implicit integrator) only if, at the stability-limited step size, the rest of the interval would take more
    // than PV_STIFF_REMAINING_STEPS steps - more than a whole RODAS4 run (300 - 400 steps at ~3x the cost each) plus the
    // latency of the extra pass costs: a few hundred rows re-run alone took 0.6 ms on top of a 3.2 ms launch when the
    // benchmark cloud's k_abs ~ 50 rows (1600 explicit steps) were handed over at a threshold of 600.
    // INT_MAX = off (an explicitly chosen DOPRI5 integrates whatever it is given).
    int stiff_after = 0x7fffffff;
    int stiff_run;     // consecutive stiff estimates (high half) | consecutive non-stiff estimates (low half)

    // z0 -> y, outputs for grid points at t0, first derivative, initial step size
    template <class Out>
    PV_HD int start(const typename Sys::Consts& ks, const double (&z0)[N], double t0, const double* tgrid, int T,
                    double rtol, double atol, Out& out) {
        st.accepted = 0;
        st.rejected = 0;
        nstep = 0;
        stiff_run = 0;
        lfacold = -13.287712f;   // log2(1e-4)
        last_rejected = false;
        t = t0;
        tend = tgrid[T - 1];
        h = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = z0[i];
        jout = 0;
        while (jout < T && tgrid[jout] <= t) {
            out(jout, tgrid[jout], y);
            ++jout;
        }
        if (jout >= T) return PV_OK;
        t_next = tgrid[jout];
        Sys::rhs(t, y, ks, k1);
        // ---- initial step (Hairer, Norsett, Wanner I, II.4) ----------------------------------
        // (fp64 here: runs once per trajectory and must survive atol -> 0 without overflow)
        double yn[N], k2[N];
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sc = atol + rtol * fabs(y[i]);
            const double a = y[i] / sc, b = k1[i] / sc;
            d0 += a * a;
            d1 += b * b;
        }
        d0 = sqrt(d0 / N);
        d1 = sqrt(d1 / N);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        h0 = fmin(fmax(h0, 1e-12), tend - t);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h0, k1[i], y[i]);
        Sys::rhs(t + h0, yn, ks, k2);
        double d2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sc = atol + rtol * fabs(y[i]);
            const double a = (k2[i] - k1[i]) / sc;
            d2 += a * a;
        }
        d2 = sqrt(d2 / N) / h0;
        const double dm = fmax(d1, d2);
        const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
        h = fmin(fmin(100.0 * h0, h1), tend - t);
        return PV_RUNNING;
    }

    // one attempted step; the 4th-order dense output of an accepted step is evaluated right away for every grid
    // point in (t, t+h] and handed to `out`
    template <class Out>
    PV_HD int step(const typename Sys::Consts& ks, const double* tgrid, int T, double rtol, double atol,
                   int max_steps, Out& out) {
        if (nstep >= max_steps) return PV_ERR_MAX_STEPS;
        ++nstep;
        bool last = false;
        if (t + 1.01 * h >= tend) {
            h = tend - t;
            last = true;
        }

        double k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], yn[N];
        // ---- stages ----------------------------------------------------------------------------
        // The newest stage derivative enters each sum LAST (outermost fma): everything else of the sum is
        // independent of it and overlaps with the preceding RHS evaluation, so the dependent chain per stage is
        // RHS -> one fma -> fma(h, sum, y) instead of one fma per stage derivative.
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h * pv_dp5[0], k1[i], y[i]);
        Sys::rhs(t + 0.2 * h, yn, ks, k2);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = fma(h, fma(pv_dp5[2], k2[i], pv_dp5[1] * k1[i]), y[i]);
        Sys::rhs(t + 0.3 * h, yn, ks, k3);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = fma(h, fma(pv_dp5[5], k3[i], fma(pv_dp5[4], k2[i], pv_dp5[3] * k1[i])), y[i]);
        Sys::rhs(t + 0.8 * h, yn, ks, k4);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = fma(h,
                        fma(pv_dp5[9], k4[i], fma(pv_dp5[8], k3[i], fma(pv_dp5[7], k2[i], pv_dp5[6] * k1[i]))),
                        y[i]);
        Sys::rhs(t + (8.0 / 9.0) * h, yn, ks, k5);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = fma(h,
                        fma(pv_dp5[14], k5[i],
                            fma(pv_dp5[13], k4[i], fma(pv_dp5[12], k3[i], fma(pv_dp5[11], k2[i], pv_dp5[10] * k1[i])))),
                        y[i]);
        Sys::rhs(t + h, yn, ks, k6);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = fma(h,
                        fma(pv_dp5[19], k6[i],
                            fma(pv_dp5[18], k5[i], fma(pv_dp5[17], k4[i], fma(pv_dp5[16], k3[i], pv_dp5[15] * k1[i])))),
                        y[i]);
        Sys::rhs(t + h, yn, ks, k7);

        // ---- error estimate (fp32 norm of an fp64 difference) -----------------------------------
        double err2d = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double e = h * fma(pv_dp5[25], k7[i],
                                     fma(pv_dp5[24], k6[i],
                                         fma(pv_dp5[23], k5[i], fma(pv_dp5[22], k4[i], fma(pv_dp5[21], k3[i], pv_dp5[20] * k1[i])))));
            const double sc = atol + rtol * pv_absmax(y[i], yn[i]);
            const double r = e * pv_rcp_approx(sc);
            err2d = fma(r, r, err2d);
        }
        float err2 = (float)(err2d * (1.0 / N));
        // a non-finite stage value makes e, hence err2, NaN or inf: treated as a failed step (give up once h is tiny)
        if (!(err2 <= 3.0e38f)) {
            err2 = 1e10f;
            if (fabs(h) < 1e-12) return PV_ERR_NONFINITE;
        }
        // Hairer's PI controller in the log2 domain (one LG2 + one EX2 per step instead of two pows and a sqrt):
        //   fac11 = err^(0.2 - 0.75 beta), beta = 0.04, err = sqrt(err2)  ->  log2 fac11 = 0.085 * log2 err2
        const float l2 = pv_log2f_fast(fmaxf(err2, 1e-30f));
        const float lfac11 = 0.085f * l2;
        if (err2 <= 1.0f) {
            // ---- accept: dense output for every grid point in (t, t+h] --------------------------
            ++st.accepted;
            const double tn = last ? tend : t + h;
            if (t_next <= tn) {
                // y(t + th h) = y + th (yd + (1 - th) (bs + th (r4 + (1 - th) r5)))   (Hairer's contd5, Horner form)
                double yd[N], bs[N], r4[N], r5[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    yd[i] = yn[i] - y[i];
                    bs[i] = fma(h, k1[i], -yd[i]);
                    r4[i] = yd[i] - fma(h, k7[i], bs[i]);
                    r5[i] = h * fma(pv_dp5[31], k7[i],
                                    fma(pv_dp5[30], k6[i],
                                        fma(pv_dp5[29], k5[i], fma(pv_dp5[28], k4[i], fma(pv_dp5[27], k3[i], pv_dp5[26] * k1[i])))));
                }
                const double hinv = pv_rcp_newton(h);
                do {
                    // a grid point at the end of the step takes th = 1 (y + yd): no branch around the interpolant
                    const double th = (t_next >= tn) ? 1.0 : (t_next - t) * hinv, th1 = 1.0 - th;
                    double yo[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) yo[i] = fma(th, fma(th1, fma(th, fma(th1, r5[i], r4[i]), bs[i]), yd[i]), y[i]);
                    out(jout, t_next, yo);
                    ++jout;
                    t_next = jout < T ? tgrid[jout] : 1e300;
                } while (t_next <= tn);
            }
            if (nstep > stiff_after && (nstep & 3) == 0) {      // every 4th step: the test costs ~30 flops + a vote
                double num = 0.0, den = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double g6 = fma(h, fma(pv_dp5[14], k5[i], fma(pv_dp5[13], k4[i], fma(pv_dp5[12], k3[i], fma(pv_dp5[11], k2[i], pv_dp5[10] * k1[i])))), y[i]);
                    const double dy = yn[i] - g6, dk = k7[i] - k6[i];
                    num = fma(dk, dk, num);
                    den = fma(dy, dy, den);
                }
                if (den > 0.0 && h * h * num > 10.5625 * den) {          // (h lambda)^2 > 3.25^2
                    stiff_run = (stiff_run & ~0xffff) + 0x10000;
                    if (stiff_run >= 15 * 0x10000 && (tend - t) > PV_STIFF_REMAINING_STEPS * h) return PV_ERR_STIFF;
                } else if (((++stiff_run) & 0xffff) >= 6) {
                    stiff_run = 0;
                }
            }
            // h_new = h / clamp(fac11 / facold^beta / 0.9, 0.1, 5)  ->  multiply by 2^-(log2 of the clamped factor)
            // (a step that follows a rejection must not grow: lower bound 0 instead of log2(1/5))
            const float lfac = fminf(fmaxf(lfac11 - 0.04f * lfacold + 0.15200309f, last_rejected ? 0.0f : -3.3219281f), 2.3219281f);
            const double hnew = h * (double)pv_exp2f_fast(-lfac);
            lfacold = fmaxf(0.5f * l2, -13.287712f);
            last_rejected = false;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                y[i] = yn[i];
                k1[i] = k7[i];
            }
            t = tn;
            if (last || jout >= T) return PV_OK;
            h = hnew;
        } else {
            ++st.rejected;
            h = h * (double)pv_exp2f_fast(-fminf(lfac11 + 0.15200309f, 2.3219281f));
            last_rejected = true;
            // only a rejection shrinks h, so the underflow test lives here and not in front of every step
            if (!(fabs(h) > 1e-14 * fmax(fabs(t), 1.0))) return PV_ERR_STEP_TOO_SMALL;
        }
        return PV_RUNNING;
    }
};

// Whole-trajectory driver (static schedule, host-side checks): integrate z from t0 over tgrid.
template <class Sys, class Out>
PV_HD int pv_dopri5_integrate(const typename Sys::Consts& ks, double (&z)[Sys::N], double t0, const double* tgrid, int T,
                              double rtol, double atol, int max_steps, Out& out, pv_step_stats& st) {
    pv_dopri5_lane<Sys> lane;
    int rc = lane.start(ks, z, t0, tgrid, T, rtol, atol, out);
    while (rc == PV_RUNNING) rc = lane.step(ks, tgrid, T, rtol, atol, max_steps, out);
    st = lane.st;
#pragma unroll
    for (int i = 0; i < Sys::N; ++i) z[i] = lane.y[i];
    return rc;
}
