// RODAS4 (Hairer & Wanner, "Solving ODEs II", IV.7: 6-stage, order 4(3), L-stable, stiffly accurate
// Rosenbrock method with a 3rd-order dense output), one trajectory per thread - the stiff path of
// north_star subsystem 2, used for the PBPK model icg_body_flat.  Replaces the per-row CVODE BDF
// solve of the reference (src/parvar/experiments/petab_factory.py:242-246) and, with SENS, the
// CVODES forward-sensitivity solve behind the objective (petab_optimization.py:144-151).
//
// Linear algebra: W = I/(gamma h) - J is assembled on the model's sparse pattern and factorised by
// the code generator's fully unrolled, symbolically ordered LU (no pivoting: for compartmental
// models -J is column diagonally dominant, so W is too).  One Jacobian, one LU, six solves and six
// RHS evaluations per step.
//
// Sensitivities: the method is applied to the augmented system (y, s_1..s_NS) with its exact, block
// lower-triangular Jacobian [[J,0],[H_k,J]], H_k = d(J s_k + f_theta_k)/dy.  Every sensitivity stage
// therefore reuses the same LU:  K_i^{s_k} = W^-1( J(y_i) s_{k,i} + f_theta_k(y_i) + H_k K_i^y + sum_j C_ij K_j^{s_k} / h ).
// The coefficient set was checked for order (4 global, 3 dense, h^4 error estimator) on a nonlinear
// problem before use - see tests/test_host_numerics.py::test_rodas4_convergence_order.
#pragma once
#include "pv_common.cuh"
#include "pv_dopri5.cuh"

namespace pv_rodas4 {
constexpr double GAMMA = 0.25;
constexpr double C2 = 0.386, C3 = 0.21, C4 = 0.63;
constexpr double A21 = 0.1544000000000000e+01;
constexpr double A31 = 0.9466785280815826e+00, A32 = 0.2557011698983284e+00;
constexpr double A41 = 0.3314825187068521e+01, A42 = 0.2896124015972201e+01, A43 = 0.9986419139977817e+00;
constexpr double A51 = 0.1221224509226641e+01, A52 = 0.6019134481288629e+01, A53 = 0.1253708332932087e+02,
                 A54 = -0.6878860361058950e+00;
constexpr double C21 = -0.5668800000000000e+01;
constexpr double C31 = -0.2430093356833875e+01, C32 = -0.2063599157091915e+00;
constexpr double C41 = -0.1073529058151375e+00, C42 = -0.9594562251023355e+01, C43 = -0.2047028614809616e+02;
constexpr double C51 = 0.7496443313967647e+01, C52 = -0.1024680431464352e+02, C53 = -0.3399990352819905e+02,
                 C54 = 0.1170890893206160e+02;
constexpr double C61 = 0.8083246795921522e+01, C62 = -0.7981132988064893e+01, C63 = -0.3152159432874371e+02,
                 C64 = 0.1631930543123136e+02, C65 = -0.6058818238834054e+01;
constexpr double D21 = 0.1012623508344586e+02, D22 = -0.7487995877610167e+01, D23 = -0.3480091861555747e+02,
                 D24 = -0.7992771707568823e+01, D25 = 0.1025137723295662e+01;
constexpr double D31 = -0.6762803392801253e+00, D32 = 0.6087714651680015e+01, D33 = 0.1643084320892478e+02,
                 D34 = 0.2476722511418386e+02, D35 = -0.6594389125716872e+01;
}  // namespace pv_rodas4

template <class M, bool SENS>
struct pv_rodas4_cost {
    static constexpr int NB = SENS ? 1 + M::NS : 1;   // blocks that go through the LU per stage
    static constexpr int FLOPS_STEP =
        M::FLOPS_JAC + M::NLU + M::FLOPS_LU + 6 * NB * M::FLOPS_SOLVE +
        6 * (SENS ? M::FLOPS_RHS_SENS + M::FLOPS_SENS_TERM + 2 * M::NS * M::NX : M::FLOPS_RHS) +
        NB * M::NX * (2 * (1 + 2 + 3 + 4 + 1 + 1) + 2 * (1 + 2 + 3 + 4 + 5) + 10) + 12;
    static constexpr int FLOPS_DENSE = NB * M::NX * (2 * 10 + 8);
};

// Per-lane RODAS4 state machine; same contract as pv_dopri5_lane.  z = (y, s_1 .. s_NS) when SENS.
template <class M, bool SENS>
struct pv_rodas4_lane {
    using Sys = pv_system<M, SENS>;
    using Consts = typename Sys::Consts;
    static constexpr int NX = M::NX;
    static constexpr int NB = SENS ? 1 + M::NS : 1;
    static constexpr int N = NX * NB;
    double t, h, tend, h_acc;
    double z[N];
    float err_acc;
    bool last_rejected;
    int jout, nstep;
    pv_step_stats st;

    // F(zs) -> r[block][state]
    PV_HD static void eval_F(double ts, const double (&zs)[N], const Consts& ks, double (&r)[NB][NX]) {
        if constexpr (SENS) {
            double y[NX], s[M::NSD][NX], f[NX], fs[M::NSD][NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) y[i] = zs[i];
#pragma unroll
            for (int p = 0; p < M::NS; ++p)
#pragma unroll
                for (int i = 0; i < NX; ++i) s[p][i] = zs[NX * (1 + p) + i];
            M::rhs_sens(ts, y, s, ks.c, ks.dc, f, fs);
#pragma unroll
            for (int i = 0; i < NX; ++i) r[0][i] = f[i];
#pragma unroll
            for (int p = 0; p < M::NS; ++p)
#pragma unroll
                for (int i = 0; i < NX; ++i) r[1 + p][i] = fs[p][i];
        } else {
            double y[NX], f[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) y[i] = zs[i];
            M::rhs(ts, y, ks.c, f);
#pragma unroll
            for (int i = 0; i < NX; ++i) r[0][i] = f[i];
        }
    }

    // solves all blocks of one stage in place: r[0] first, then r[k] += H_k r[0], solve (same LU)
    PV_HD void solve_stage(const Consts& ks, const double (&W)[M::NLU], double (&r)[NB][NX]) const {
        M::lu_solve(W, r[0]);
        if constexpr (SENS) {
            double yn0[NX], sn0[M::NSD][NX], hk[M::NSD][NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) yn0[i] = z[i];
#pragma unroll
            for (int p = 0; p < M::NS; ++p)
#pragma unroll
                for (int i = 0; i < NX; ++i) sn0[p][i] = z[NX * (1 + p) + i];
            M::sens_stage_term(t, yn0, sn0, r[0], ks.c, ks.dc, hk);
#pragma unroll
            for (int p = 0; p < M::NS; ++p) {
#pragma unroll
                for (int i = 0; i < NX; ++i) r[1 + p][i] += hk[p][i];
                M::lu_solve(W, r[1 + p]);
            }
        }
    }

    template <class Out>
    PV_HD int start(const Consts& ks, const double (&z0)[N], double t0, const double* tgrid, int T, double rtol,
                    double atol, Out& out) {
        st.accepted = 0;
        st.rejected = 0;
        nstep = 0;
        last_rejected = false;
        err_acc = 1e-2f;
        t = t0;
        tend = tgrid[T - 1];
        h = 0.0;
        h_acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = z0[i];
        jout = 0;
        while (jout < T && tgrid[jout] <= t) {
            out(jout, tgrid[jout], z);
            ++jout;
        }
        if (jout >= T) return PV_OK;
        // ---- initial step size ---------------------------------------------------------------------
        double F0[NB][NX];
        eval_F(t, z, ks, F0);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const float sc = (float)(atol + rtol * fabs(z[b * NX + i]));
                const float a = (float)z[b * NX + i] / sc, q = (float)F0[b][i] / sc;
                d0 += a * a;
                d1 += q * q;
            }
        d0 = sqrtf(d0 / N);
        d1 = sqrtf(d1 / N);
        h = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6 : 0.01 * (double)(d0 / d1);
        h = fmin(fmax(h, 1e-8), tend - t);
        h_acc = h;
        return PV_RUNNING;
    }

    template <class Out>
    PV_HD int step(const Consts& ks, const double* tgrid, int T, double rtol, double atol, int max_steps, Out& out) {
        using namespace pv_rodas4;
        if (nstep >= max_steps) return PV_ERR_MAX_STEPS;
        ++nstep;
        bool last = false;
        if (t + 1.01 * h >= tend) {
            h = tend - t;
            last = true;
        }
        if (!(fabs(h) > 1e-14 * fmax(fabs(t), 1.0))) return PV_ERR_STEP_TOO_SMALL;
        const double hi = pv_rcp_newton(h);   // h is a normal number here (checked above): no IEEE slow path

        double K1[NB][NX], K2[NB][NX], K3[NB][NX], K4[NB][NX], K5[NB][NX], K6[NB][NX];
        double zs[N];
        double W[M::NLU];
        // ---- Jacobian, W = I/(gamma h) - J, LU ------------------------------------------------------------
        {
            double J[M::NNZ], yn0[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) yn0[i] = z[i];
            M::jac(t, yn0, ks.c, J);
            M::lu_assemble(hi * (1.0 / GAMMA), J, W);
            M::lu_factor(W);
        }
        // (in every stage sum below the newest K enters last, as the outermost fma: the rest of the sum does not wait for it)
        // ---- stage 1 ----------------------------------------------------------------------------------
        eval_F(t, z, ks, K1);
        solve_stage(ks, W, K1);
        // ---- stage 2 ----------------------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) zs[b * NX + i] = fma(A21, K1[b][i], z[b * NX + i]);
        eval_F(t + C2 * h, zs, ks, K2);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) K2[b][i] = fma(hi * C21, K1[b][i], K2[b][i]);
        solve_stage(ks, W, K2);
        // ---- stage 3 ----------------------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) zs[b * NX + i] = fma(A32, K2[b][i], fma(A31, K1[b][i], z[b * NX + i]));
        eval_F(t + C3 * h, zs, ks, K3);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) K3[b][i] = fma(hi, fma(C32, K2[b][i], C31 * K1[b][i]), K3[b][i]);
        solve_stage(ks, W, K3);
        // ---- stage 4 ----------------------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i)
                zs[b * NX + i] = fma(A43, K3[b][i], fma(A42, K2[b][i], fma(A41, K1[b][i], z[b * NX + i])));
        eval_F(t + C4 * h, zs, ks, K4);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i)
                K4[b][i] = fma(hi, fma(C43, K3[b][i], fma(C42, K2[b][i], C41 * K1[b][i])), K4[b][i]);
        solve_stage(ks, W, K4);
        // ---- stage 5 ----------------------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i)
                zs[b * NX + i] =
                    fma(A54, K4[b][i], fma(A53, K3[b][i], fma(A52, K2[b][i], fma(A51, K1[b][i], z[b * NX + i]))));
        eval_F(t + h, zs, ks, K5);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i)
                K5[b][i] =
                    fma(hi, fma(C54, K4[b][i], fma(C53, K3[b][i], fma(C52, K2[b][i], C51 * K1[b][i]))), K5[b][i]);
        solve_stage(ks, W, K5);
        // ---- stage 6 (the embedded solution zs + K5 is the evaluation point) -------------------------------
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) zs[b * NX + i] += K5[b][i];
        eval_F(t + h, zs, ks, K6);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i)
                K6[b][i] = fma(hi,
                               fma(C65, K5[b][i],
                                   fma(C64, K4[b][i], fma(C63, K3[b][i], fma(C62, K2[b][i], C61 * K1[b][i])))),
                               K6[b][i]);
        solve_stage(ks, W, K6);
        // new solution = zs + K6, error estimate = K6
        // error norm with a ~20-bit reciprocal of the scale (one MUFU.RCP64H, no division slow path): it only steers h
        double err2d = 0.0;
        bool finite = true;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double znew = zs[b * NX + i] + K6[b][i];
                const double sc = fma(rtol, pv_absmax(z[b * NX + i], znew), atol);
                const double r = K6[b][i] * pv_rcp_approx(sc);
                err2d = fma(r, r, err2d);
                finite = finite && (fabs(znew) < 1e300);
                zs[b * NX + i] = znew;
            }
        float err2 = (float)(err2d * (1.0 / N));
        if (!finite || !(err2 <= 3.0e38f)) {
            err2 = 1e10f;
            if (fabs(h) < 1e-12) return PV_ERR_NONFINITE;
        }
        // hnew = h / fac, fac = err^(1/4) / safe clipped to [1/6, 5]
        const float err4 = pv_powf_fast(fmaxf(err2, 1e-30f), 0.125f);
        float fac = fmaxf(1.0f / 6.0f, fminf(5.0f, err4 * (1.0f / 0.9f)));
        if (err2 <= 1.0f) {
            ++st.accepted;
            const double tn = last ? tend : t + h;
            if (jout < T && tgrid[jout] <= tn) {
                do {
                    const double tj = tgrid[jout];
                    double zo[N];
                    if (tj >= tn) {
#pragma unroll
                        for (int i = 0; i < N; ++i) zo[i] = zs[i];
                    } else {
                        const double s = (tj - t) * hi, s1 = 1.0 - s;
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int i = 0; i < NX; ++i) {
                                const double c2 = fma(D21, K1[b][i], fma(D22, K2[b][i], fma(D23, K3[b][i], fma(D24, K4[b][i], D25 * K5[b][i]))));
                                const double c3 = fma(D31, K1[b][i], fma(D32, K2[b][i], fma(D33, K3[b][i], fma(D34, K4[b][i], D35 * K5[b][i]))));
                                zo[b * NX + i] = fma(s1, z[b * NX + i], s * fma(s1, fma(s, c3, c2), zs[b * NX + i]));
                            }
                    }
                    out(jout, tj, zo);
                    ++jout;
                } while (jout < T && tgrid[jout] <= tn);
            }
            // Gustafsson predictive controller (rodas.f, PRED)
            if (st.accepted > 1) {
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
            for (int i = 0; i < N; ++i) z[i] = zs[i];
            t = tn;
            if (last || jout >= T) return PV_OK;
            h = hnew;
        } else {
            ++st.rejected;
            h = (st.accepted == 0 && st.rejected >= 2) ? h * 0.1 : h * (double)pv_fdiv_fast(1.0f, fac);
            last_rejected = true;
        }
        return PV_RUNNING;
    }
};

template <class M, bool SENS, class Out>
PV_HD int pv_rosenbrock_integrate(const typename pv_system<M, SENS>::Consts& ks,
                                  double (&z)[pv_system<M, SENS>::N], double t0, const double* tgrid, int T, double rtol,
                                  double atol, int max_steps, Out& out, pv_step_stats& st) {
    pv_rodas4_lane<M, SENS> lane;
    int rc = lane.start(ks, z, t0, tgrid, T, rtol, atol, out);
    while (rc == PV_RUNNING) rc = lane.step(ks, tgrid, T, rtol, atol, max_steps, out);
    st = lane.st;
#pragma unroll
    for (int i = 0; i < pv_system<M, SENS>::N; ++i) z[i] = lane.z[i];
    return rc;
}
