// DOPRI5 with forward sensitivities for SMALL batches, one trajectory per GROUP of NS lanes (BASELINE config 3 as it is
// literally stated: 8 chains x 1000 individuals = 8000 trajectories).  Such a launch is a single wave, so it lasts as long
// as its longest trajectory's DEPENDENT CHAIN - 668 steps here - at one warp per scheduler; throughput does not matter,
// cycles per step do.  Measured on B200 (tools/probe_latency.cu): a dependent DFMA costs 8 cycles, moving a double to
// another lane 26, an F2F 19.  The first small-batch kernel (state block and sensitivity blocks on 1 + NS lanes, removed;
// ncu summary profiles/r02_c3_roles_ncu_full.txt) paid a 26-cycle shuffle in front of each of the seven stage evaluations,
// role selects on every operand and, with 10 trajectories per warp, 1.35 warps per scheduler under a 128-register cap with
// spills: 1650 cycles per step.  Here lane k of a group integrates the STATE block itself plus sensitivity block k
// (1380 cycles per step, profiles/r02_c3_pairs_ncu_full.txt):
//   * nothing crosses lanes inside a step except the error norm (one shuffle per further lane of the group + the warp's REDUX):
//     the state chain and the sensitivity chain of a lane are independent instruction streams that overlap;
//   * 32 / NS trajectories per warp (16 for two parameters): 8000 trajectories are 500 warps <= 592 schedulers, launched as
//     one-warp CTAs so they spread over all SMs, and no register cap (no spills);
//   * the state block is integrated NS times (redundant flops on an fp64 pipe that is 30 % busy in this regime).
// The step itself is pv_dopri5_warp (one step size per warp); the trajectory's error norm is the same mean square over all
// NX (1 + NS) components as in the other kernels (state block counted once).  Results agree with them within the
// integration tolerance.  Replaces, like them, the AMICI forward-sensitivity solve behind pyPESTO's objective
// (petab_optimization.py:144-151).
#pragma once
#include "pv_dopri5_warp.cuh"

#define PV_PAIRS_BLOCK 32

template <class M>
struct pv_pair_system {
    static constexpr int NX = M::NX;
    static constexpr int N = 2 * M::NX;
    static constexpr int GROUP = M::NS;          // lanes per trajectory
    static constexpr int ERR_SHARED = M::NX;     // leading components every lane of the group carries (counted once)
    static constexpr int ERR_N = M::NX * (1 + M::NS);
    struct Consts {
        double c[M::NC];
        double w[M::NC];                         // d c / d theta_k of this lane's parameter, dense
    };
    PV_HD static void rhs(double t, const double (&z)[N], const Consts& k, double (&fz)[N]) {
        double y[NX], v[NX], f[NX], o[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            y[i] = z[i];
            v[i] = z[NX + i];
        }
        M::rhs(t, y, k.c, f);
        M::jv_plus_dfdc(t, y, v, k.c, k.w, o);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            fz[i] = f[i];
            fz[NX + i] = o[i];
        }
    }
};

// per-grid-point sink of a lane: logp (every lane of the group accumulates it, lane 0 reports it) and d logp / d theta_k
template <class M, bool QSH>
struct pv_pair_sink {
    static constexpr int NX = M::NX;
    const pv_launch_params& p;
    const double* qrow;
    unsigned qaddr;
    double obs[NX];
    double lp, g;
    PV_D void operator()(int, double, const double (&z)[2 * NX]) {
        for (int k = 0; k < p.n_obs; ++k) {
            const int idx = p.obs_idx[k];
            double sc = 0.0, f = 0.0, df = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                sc = idx == i ? obs[i] : sc;
                f = idx == i ? z[i] : f;
                df = idx == i ? z[NX + i] : df;
            }
            f *= sc;
            df *= sc;
            double q0, q1, q2;
            if constexpr (QSH) {
#if defined(__CUDA_ARCH__)
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q0), "=d"(q1) : "r"(qaddr + k * (PV_QSTRIDE * 8)));
                asm("ld.shared.f64 %0, [%1];" : "=d"(q2) : "r"(qaddr + k * (PV_QSTRIDE * 8) + 16));
#else
                q0 = q1 = q2 = 0.0;
#endif
            } else {
                const double2 q01 = *reinterpret_cast<const double2*>(qrow + k * PV_QSTRIDE);
                q2 = qrow[k * PV_QSTRIDE + 2];
                q0 = q01.x;
                q1 = q01.y;
            }
            if (p.noise_model == 0) {
                lp += fma(f, fma(q2, f, q1), q0);
                g = fma(fma(2.0 * q2, f, q1), df, g);
            } else if (q2 != 0.0) {
                const double lf = log(f);
                lp += fma(lf, fma(q2, lf, q1), q0);
                g += fma(2.0 * q2, lf, q1) * df / f;
            }
        }
        if constexpr (QSH)
            qaddr += p.n_obs * (PV_QSTRIDE * 8);
        else
            qrow += p.n_obs * PV_QSTRIDE;
    }
};

template <class M, bool QSH>
__global__ void __launch_bounds__(PV_PAIRS_BLOCK)
pv_batch_pairs_kernel(const __grid_constant__ pv_launch_params_m<M> p) {
    constexpr int NX = M::NX, NS = M::NS, G = NS, TPW = 32 / G;
    static_assert(NS >= 1 && G <= 32, "needs at least one sensitivity parameter");
    using Sys = pv_pair_system<M>;
    using Lane = pv_dopri5_warp<Sys>;
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
    // lanes beyond the last whole group repeat the last group's lane 0 (identical values to identical addresses)
    const int tr = min(lane_id / G, TPW - 1);
    const int role = lane_id < TPW * G ? lane_id - tr * G : 0;      // this lane's sensitivity parameter
    const int64_t n_items = p.n_items_dev ? (int64_t)*p.n_items_dev : p.B;
    typename Sys::Consts ks;
    Lane lane;
    lane.stiff_after = p.stiff_after;
    lane.group_src = tr * G;
    pv_pair_sink<M, QSH> sink{p, nullptr, 0u, {}, 0.0, 0.0};

    for (;;) {
        unsigned long long base = 0;
        if (lane_id == 0) base = atomicAdd(p.counter, (unsigned long long)TPW);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= n_items) break;
        // tiles from the END of the work list (the high-key end under the locality schedule: the longest first)
        const int64_t tile0 = (n_items - 1) / TPW * TPW - (int64_t)base;
        const int64_t item = min(tile0 + tr, n_items - 1);
        const int64_t b = p.index_map ? (int64_t)p.index_map[item] : item;

        double z[Sys::N];
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
            M::hoist(th, ks.c, y0, sink.obs);
            M::hoist_sens(th, dc, dy0);
            M::dc_dense(role, dc, ks.w);
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
                for (int s = 0; s < NS; ++s) s0 = (role == s && !over) ? dy0[s][i] : s0;
                z[i] = x;
                z[NX + i] = s0;
            }
        }
        sink.lp = 0.0;
        sink.g = 0.0;
        if constexpr (QSH) {
            sink.qaddr = (unsigned)__cvta_generic_to_shared(sh_stats);
        } else {
            sink.qrow = p.stats + (int64_t)(b % p.n_data) * (p.T * p.n_obs * PV_QSTRIDE);
        }
        int rc = lane.start(ks, z, p.t0, sh_t, p.T, p.rtol, p.atol, sink);
        while (rc == PV_RUNNING)
            rc = lane.cur ? lane.template step<1>(ks, sh_t, p.T, p.rtol, p.atol, p.max_steps, sink)
                          : lane.template step<0>(ks, sh_t, p.T, p.rtol, p.atol, p.max_steps, sink);
        // a trajectory is dead if any lane of its group is (every lane takes part in the shuffles: no short-circuit)
        const double n_dead = lane.group_sum(lane.dead ? 1.0 : 0.0);
        if (n_dead > 0.0) rc = PV_ERR_NONFINITE;

        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        p.grad[(int64_t)role * p.B + b] = (rc == 0) ? sink.g : nan;
        if (role == 0) {
            p.logp[b] = (rc == 0) ? sink.lp : nan;
            if (p.status) p.status[b] = rc;
            if (p.steps) {
                p.steps[b] = lane.st.accepted;
                p.steps[p.B + b] = lane.st.rejected;
            }
        }
    }
}
