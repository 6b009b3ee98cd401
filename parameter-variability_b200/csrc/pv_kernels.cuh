// Batched trajectory / likelihood kernels: one trajectory per thread, everything per-lane in
// registers, time grid + shared data in shared memory, batch-minor (coalescing) global layout.
// Replaces the reference's per-row roadrunner loop (petab_factory.py:228-246) and the AMICI
// objective (petab_optimization.py:144-151, 168-173).  See include/parvar_b200.h for the contract.
#pragma once
#include <cuda_runtime.h>
#include "pv_common.cuh"
#include "pv_dopri5.cuh"
#include "pv_dopri5_warp.cuh"
#include "pv_rosenbrock.cuh"

#define PV_MAX_COLS 8
#define PV_MAX_OBS 16
#ifndef PV_BLOCK
#define PV_BLOCK 128
#endif
#ifndef PV_MINB_SMALL
#define PV_MINB_SMALL 4
#endif
#ifndef PV_MINB_STIFF
#define PV_MINB_STIFF 1
#endif

// mode = bit set: what a launch produces
enum pv_mode : int {
    PV_MODE_TRAJ = 1,        // write Y (and S when SENS)
    PV_MODE_LOGP = 2,        // accumulate the log-likelihood (and its gradient when SENS)
    PV_MODE_SENS = 4,        // integrate the forward-sensitivity system as well
    PV_MODE_TRAJ_LOGP = 3,
    PV_MODE_LOGP_GRAD = 6,
    PV_MODE_TRAJ_SENS = 5,
    PV_MODE_FIM = 8,         // also accumulate the Fisher information (needs LOGP | SENS)
    PV_MODE_LOGP_GRAD_FIM = 14,
};

struct pv_launch_params {
    int64_t B;
    const double* P;           // [n_cols][B]
    const double* tgrid;       // [T]
    const double* stats;       // [n_data][T][n_obs][4]  likelihood polynomial coefficients q0, q1, q2, pad (see pv_problem_set_data)
    double* Y;                 // [B][T][n_obs]
    double* S;                 // [B][T][n_obs][NS]
    double* logp;              // [B]
    double* grad;              // [NS][B]
    double* fim;               // [NS(NS+1)/2][B]  upper triangle, row-major
    int32_t* status;           // [B]
    int32_t* steps;            // [2][B]
    double rtol, atol;
    int col_target[PV_MAX_COLS];   // < NTH: theta index; >= NTH: NTH + state index (initial value)
    int obs_idx[PV_MAX_OBS];
    unsigned long long* counter;   // next unprocessed work item (lane refill)
    const int32_t* index_map;      // work item -> trajectory (nullptr: identity); used by the stiff re-run pass
    const int* n_items_dev;        // number of work items if it lives on the device (nullptr: B)
    double t0;                     // integration start (initial values hold at t0 <= tgrid[0])
    int n_cols, T, n_obs, n_data, noise_model, max_steps, refill_min;
    int stiff_after;               // DOPRI5 stiffness detection starts after this many steps (INT_MAX = off)
};

// The batch-wide values (`setValue` calls shared by all rows) ride in the kernel parameters themselves: a launch
// carries its own snapshot, so changing a value between launches cannot race with a kernel still in flight on another
// stream, and no host->device copy of configuration data sits in front of a launch.
template <class M>
struct pv_launch_params_m : pv_launch_params {
    double theta_base[M::NTH];   // batch-wide theta
    double y0_base[M::NX];       // NaN = use the model's initial value
};

template <int NX, int N>
PV_D double pv_pick(const double (&a)[N], int block, int idx) {
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) v = (idx == i) ? a[block * NX + i] : v;
    return v;
}

// ---- per-grid-point sink: stores the observables and / or folds them into the likelihood ---------------------
// The integrators call it once per grid point in increasing j, so a lane walks three running pointers (its row of Y,
// of S and of the likelihood coefficients) instead of rebuilding addresses from (b, j, k) at every point.
// OBSID = true: the observables are exactly the model's states in order (n_obs == NX, obs_idx[k] == k), so the
// selection chains, the indexed constant loads and the k-loop overhead disappear (the common case of the
// cheap non-stiff models, where the per-output code is a sizeable share of a step).
#define PV_QSTRIDE 4   // doubles per likelihood coefficient record (q0, q1, q2, pad): two 16-byte loads
// QSH = true: the coefficient records are known to sit in shared memory; they are read with ld.shared through a 32-bit
// address (the generic-address loads of the default path may alias the trajectory stores as far as the compiler knows,
// which pins them behind the stores and puts their latency on the critical path of every grid point).
// STAGE = true (lockstep warps, trajectories without sensitivities): the observables of PV_STAGE_CAP / n_obs grid points
// are collected in a per-warp shared-memory block [32 rows][PV_STAGE_RS] and written out by the whole warp, consecutive
// lanes to consecutive doubles of one trajectory's row.  A lane storing its own 8 bytes per observable and grid point
// costs one L1 wavefront per lane and store instruction (32 rows = 32 lines): 150 stores x 32 wavefronts per tile of 32
// trajectories, 0.65 ms of the 3.03 ms C2 launch (measured: the same launch without the trajectory store, 2.39 ms).
// Transposed, a store instruction covers 256 contiguous bytes of ~1.3 rows (3-4 wavefronts), and the rows fill whole
// 32-byte sectors at once instead of 8 bytes at a time.
#ifndef PV_STAGE_CAP
#define PV_STAGE_CAP 24                 // doubles staged per trajectory between flushes
#endif
#if defined(PV_STAGE_CS)
#define PV_STAGE_STORE(ptr, v) __stcs(ptr, v)
#else
#define PV_STAGE_STORE(ptr, v) (*(ptr) = (v))
#endif
#ifndef PV_STAGE_ON
#define PV_STAGE_ON 1
#endif
#define PV_STAGE_RS (PV_STAGE_CAP + 1)  // row stride (odd: conflict-free 8-byte accesses); the last slot holds the row's address
template <class M, int MODE, bool OBSID, bool QSH = false, bool STAGE = false>
struct pv_sink {
    static constexpr bool SENS = (MODE & PV_MODE_SENS) != 0;
    static constexpr bool TRAJ = (MODE & PV_MODE_TRAJ) != 0;
    static constexpr bool LOGP = (MODE & PV_MODE_LOGP) != 0;
    static constexpr bool FIM = (MODE & PV_MODE_FIM) != 0;
    static constexpr int N = pv_system<M, SENS>::N;
    static constexpr int NFIM = FIM ? M::NSD * (M::NSD + 1) / 2 : 1;
    const pv_launch_params& p;
    int64_t b;                // trajectory
    double* yrow;             // &Y[b][j][0] of the next grid point (trajectory-major: see below)
    double* srow;             // &S[b][j][0][0]
    const double* qrow;       // &q[d][j][0][0]: coefficients of the trajectory's data set d = b % n_data, either in the
                              // block's shared-memory copy (n_data == 1) or in global memory (generic address)
    unsigned qaddr = 0;       // the same as a shared-memory address (QSH)
    unsigned st_base = 0;     // STAGE: shared-memory address of the warp's staging block
    unsigned st_w = 0;        //        of this lane's next free slot
    int st_cnt = 0;           //        doubles staged per lane (warp-uniform)
    int st_off = 0;           //        offset inside the trajectory's row of the first staged double
    double obs[M::NX];
    double lp;
    double g[M::NSD];
    double fim[NFIM];         // sum over measurements of (df/dtheta)(df/dtheta)^T / sigma^2  (AMICI's FIM)

    // first grid point of trajectory b_
    PV_D void begin(int64_t b_, const double* qbase) {
        const int n_obs = OBSID ? M::NX : p.n_obs;
        b = b_;
        if constexpr (TRAJ) {
            const int64_t off = b_ * p.T * n_obs;
            yrow = p.Y + off;
            if constexpr (SENS) srow = p.S + off * M::NS;
            if constexpr (STAGE) {
#if defined(__CUDA_ARCH__)
                st_w = st_base + (threadIdx.x & 31) * (PV_STAGE_RS * 8);
                asm volatile("st.shared.u64 [%0], %1;" ::"r"(st_w + PV_STAGE_CAP * 8), "l"(yrow) : "memory");
#endif
                st_cnt = 0;
                st_off = 0;
            }
        }
        if constexpr (LOGP) {
            if constexpr (QSH) {
#if defined(__CUDA_ARCH__)
                qaddr = (unsigned)__cvta_generic_to_shared(qbase);      // n_data == 1: every trajectory reads the one record set
#endif
            } else {
                qrow = qbase + (int64_t)(b_ % p.n_data) * (p.T * n_obs * PV_QSTRIDE);
            }
        }
        lp = 0.0;
#pragma unroll
        for (int s = 0; s < M::NSD; ++s) g[s] = 0.0;
#pragma unroll
        for (int q = 0; q < NFIM; ++q) fim[q] = 0.0;
    }

    template <bool LOGNORMAL>
    PV_D void one(int k, double f, const double (&df)[M::NSD]) {
        if constexpr (TRAJ) {
            // trajectory-major: a lane writes its own contiguous [T][n_obs] block, so consecutive grid points of one
            // trajectory fill whole 32-byte sectors and 128-byte lines no matter which lane integrates it (lane refill
            // scatters neighbouring trajectories over warps and time; batch-minor output made every store a partial
            // sector that was evicted half filled: 2.7x the algorithmic DRAM traffic in ncu)
            if constexpr (STAGE) {
#if defined(__CUDA_ARCH__)
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(st_w + k * 8), "d"(f) : "memory");
#endif
            } else {
                yrow[k] = f;
            }
            if constexpr (SENS) {
#pragma unroll
                for (int s = 0; s < M::NS; ++s) srow[k * M::NS + s] = df[s];
            }
        }
        if constexpr (LOGP) {
            // lp term = q0 + v (q1 + q2 v), v = f (normal) or log f (log-normal); n = 0 entries have q = 0
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
            if constexpr (!LOGNORMAL) {
                lp += fma(f, fma(q2, f, q1), q0);
                if constexpr (SENS) {
                    const double w = fma(2.0 * q2, f, q1);
#pragma unroll
                    for (int s = 0; s < M::NS; ++s) g[s] = fma(w, df[s], g[s]);
                    if constexpr (FIM) {
                        const double wi = -2.0 * q2;
                        int q = 0;
#pragma unroll
                        for (int s = 0; s < M::NS; ++s)
#pragma unroll
                            for (int r = s; r < M::NS; ++r) {
                                fim[q] = fma(wi * df[s], df[r], fim[q]);
                                ++q;
                            }
                    }
                }
            } else if (q2 != 0.0) {
                const double lf = log(f);
                lp += fma(lf, fma(q2, lf, q1), q0);
                if constexpr (SENS) {
                    const double rf = 1.0 / f;
                    const double w = fma(2.0 * q2, lf, q1) * rf;
#pragma unroll
                    for (int s = 0; s < M::NS; ++s) g[s] = fma(w, df[s], g[s]);
                    if constexpr (FIM) {
                        const double wi = -2.0 * q2 * rf * rf;
                        int q = 0;
#pragma unroll
                        for (int s = 0; s < M::NS; ++s)
#pragma unroll
                            for (int r = s; r < M::NS; ++r) {
                                fim[q] = fma(wi * df[s], df[r], fim[q]);
                                ++q;
                            }
                    }
                }
            }
        }
    }

    template <bool LOGNORMAL>
    PV_D void row(const double (&z)[N]) {
        if constexpr (OBSID) {
#pragma unroll
            for (int k = 0; k < M::NX; ++k) {
                double df[M::NSD];
                if constexpr (SENS) {
#pragma unroll
                    for (int s = 0; s < M::NS; ++s) df[s] = M::OBS_UNIT ? z[M::NX * (1 + s) + k] : z[M::NX * (1 + s) + k] * obs[k];
                }
                one<LOGNORMAL>(k, M::OBS_UNIT ? z[k] : z[k] * obs[k], df);
            }
        } else {
            for (int k = 0; k < p.n_obs; ++k) {
                const int idx = p.obs_idx[k];
                const double sc = pv_pick<M::NX>(obs, 0, idx);
                double df[M::NSD];
                if constexpr (SENS) {
#pragma unroll
                    for (int s = 0; s < M::NS; ++s) df[s] = pv_pick<M::NX>(z, 1 + s, idx) * sc;
                }
                one<LOGNORMAL>(k, pv_pick<M::NX>(z, 0, idx) * sc, df);
            }
        }
    }

    PV_D void operator()(int, double, const double (&z)[N]) {
        // the noise model is a launch-wide constant: one uniform branch per grid point, not one per observable
        if (LOGP && p.noise_model != 0)
            row<true>(z);
        else
            row<false>(z);
        const int n_obs = OBSID ? M::NX : p.n_obs;
        if constexpr (TRAJ) {
            if constexpr (STAGE) {
                st_w += n_obs * 8;
                st_cnt += n_obs;
                if (st_cnt + n_obs > PV_STAGE_CAP) flush();       // warp-uniform
            } else {
                yrow += n_obs;
            }
            if constexpr (SENS) srow += n_obs * M::NS;
        }
        if constexpr (LOGP) {
            if constexpr (QSH)
                qaddr += n_obs * (PV_QSTRIDE * 8);
            else
                qrow += n_obs * PV_QSTRIDE;
        }
    }

    // STAGE: the warp writes the st_cnt staged doubles of each of its 32 rows; flat index f = row * CNT + k over the lanes
    template <int CNT>
    PV_D void flush_rows(int cnt_rt) {
#if defined(__CUDA_ARCH__)
        const int cnt = CNT > 0 ? CNT : cnt_rt;
        const int lane_id = threadIdx.x & 31;
#pragma unroll 4
        for (int i = 0; i < cnt; ++i) {
            const int f = i * 32 + lane_id;
            const int r = f / cnt, k = f - r * cnt;
            const unsigned a = st_base + r * (PV_STAGE_RS * 8);
            double v;
            unsigned long long row;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + k * 8) : "memory");
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(row) : "r"(a + PV_STAGE_CAP * 8) : "memory");
            PV_STAGE_STORE(reinterpret_cast<double*>(row) + st_off + k, v);
        }
#endif
    }
    PV_D void flush() {
        if constexpr (STAGE) {
            if (st_cnt == 0) return;
            constexpr int FULL_CNT = OBSID ? (PV_STAGE_CAP / M::NX) * M::NX : 0;
            __syncwarp();
            if (FULL_CNT > 0 && st_cnt == FULL_CNT)
                flush_rows<FULL_CNT>(FULL_CNT);
            else
                flush_rows<0>(st_cnt);
            __syncwarp();
            st_off += st_cnt;
            st_cnt = 0;
#if defined(__CUDA_ARCH__)
            st_w = st_base + (threadIdx.x & 31) * (PV_STAGE_RS * 8);
#endif
        }
    }
};

// ---- integrator selection ---------------------------------------------------------------------------------
template <class M, int METHOD, bool SENS>
struct pv_lane_of {
    using type = pv_dopri5_lane<pv_system<M, SENS>>;
};
template <class M, bool SENS>
struct pv_lane_of<M, 2, SENS> {
    using type = pv_rodas4_lane<M, SENS>;
};

// ---- the kernel ----------------------------------------------------------------------------------------
// Persistent warps with LANE REFILL: a lane that finishes its trajectory takes the next unprocessed one
// (one warp-aggregated atomicAdd on a global counter) while its 31 neighbours keep stepping, so a warp
// never idles behind its slowest trajectory.  Step counts over a log-normal parameter cloud differ by
// ~10x (SURVEY.md section 7 "Divergence"); with one fixed trajectory per lane the warp pays max-over-lanes.
// Refills are batched (>= p.refill_min idle lanes, or nobody left working) so the start-up code
// (parameter load, hoist, initial step size) runs with several lanes active.  Results do not depend on
// the schedule: every trajectory is integrated by exactly one lane from its own inputs.
// resident CTAs per SM the register allocation must allow: 4 (128 registers) for the small explicit systems, where
// occupancy hides the fp64 dependency latency; the augmented / Rosenbrock kernels take what they need
template <class M, int METHOD, int MODE>
struct pv_kernel_traits {
    static constexpr bool SENS = (MODE & PV_MODE_SENS) != 0;
    static constexpr int MIN_BLOCKS = (METHOD == 1 && pv_system<M, SENS>::N <= 4) ? PV_MINB_SMALL : (METHOD == 2 ? PV_MINB_STIFF : 1);
};

template <class M, int METHOD, int MODE, bool OBSID>
__global__ void __launch_bounds__(PV_BLOCK, pv_kernel_traits<M, METHOD, MODE>::MIN_BLOCKS)
pv_batch_kernel(const __grid_constant__ pv_launch_params_m<M> p) {
    constexpr bool SENS = (MODE & PV_MODE_SENS) != 0;
    constexpr bool LOGP = (MODE & PV_MODE_LOGP) != 0;
    using Sys = pv_system<M, SENS>;
    using Lane = typename pv_lane_of<M, METHOD, SENS>::type;
    extern __shared__ __align__(16) double sh[];
    // the coefficient records come first: they are read with 16-byte loads, and their size is a multiple of 32 bytes
    const bool stats_shared = LOGP && p.n_data == 1;
    double* sh_stats = sh;                                                    // [T][n_obs][4] when stats_shared
    double* sh_t = sh + (stats_shared ? PV_QSTRIDE * p.T * p.n_obs : 0);      // [T]
    double* sh_theta = sh_t + p.T;                                            // [NTH]
    double* sh_y0 = sh_theta + M::NTH;                                        // [NX]
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) sh_t[i] = p.tgrid[i];
    for (int i = threadIdx.x; i < M::NTH; i += blockDim.x) sh_theta[i] = p.theta_base[i];
    for (int i = threadIdx.x; i < M::NX; i += blockDim.x) sh_y0[i] = p.y0_base[i];
    if (stats_shared)
        for (int i = threadIdx.x; i < PV_QSTRIDE * p.T * p.n_obs; i += blockDim.x) sh_stats[i] = p.stats[i];
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    const int64_t n_items = p.n_items_dev ? (int64_t)*p.n_items_dev : p.B;
    typename Sys::Consts ks;
    Lane lane;
    pv_sink<M, MODE, OBSID> sink{p, 0, nullptr, nullptr, nullptr};
    const double* qbase = stats_shared ? sh_stats : p.stats;
    bool have = false;        // this lane is integrating trajectory sink.b
    bool exhausted = false;   // the work counter ran past B

    auto finish = [&](int rc) {
        const int64_t b = sink.b;
        if constexpr (LOGP) {
            // a trajectory that did not finish has no likelihood: NaN in logp, gradient and Fisher information alike
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            p.logp[b] = (rc == 0) ? sink.lp : nan;
            if constexpr (SENS) {
#pragma unroll
                for (int s = 0; s < M::NS; ++s) p.grad[(int64_t)s * p.B + b] = (rc == 0) ? sink.g[s] : nan;
                if constexpr ((MODE & PV_MODE_FIM) != 0) {
#pragma unroll
                    for (int q = 0; q < M::NS * (M::NS + 1) / 2; ++q) p.fim[(int64_t)q * p.B + b] = (rc == 0) ? sink.fim[q] : nan;
                }
            }
        }
        if (p.status) p.status[b] = rc;
        if (p.steps) {
            p.steps[b] = lane.st.accepted;
            p.steps[p.B + b] = lane.st.rejected;
        }
        have = false;
    };

    while (true) {
        __syncwarp();
        const unsigned need = __ballot_sync(FULL, !have && !exhausted);
        const unsigned busy = __ballot_sync(FULL, have);
        if (need && (__popc(need) >= p.refill_min || busy == 0)) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if (lane_id == leader) base = atomicAdd(p.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if ((need >> lane_id) & 1u) {
                const int64_t taken = (int64_t)base + __popc(need & ((1u << lane_id) - 1u));
                // handed out from the END of the work list: the high-key end under the locality schedule, where the
                // longest trajectories are - they start first instead of forming the launch's tail
                const int64_t item = n_items - 1 - taken;
                if (taken >= n_items) {
                    exhausted = true;
                } else {
                    const int64_t b = p.index_map ? (int64_t)p.index_map[item] : item;
                    // ---- theta = batch-wide values, overridden by this trajectory's columns -----------------
                    double pc[PV_MAX_COLS];
#pragma unroll
                    for (int j = 0; j < PV_MAX_COLS; ++j) pc[j] = (j < p.n_cols) ? __ldg(p.P + (int64_t)j * p.B + b) : 0.0;
                    double z[Sys::N];
                    double th[M::NTH];
#pragma unroll
                    for (int i = 0; i < M::NTH; ++i) {
                        double v = sh_theta[i];
#pragma unroll
                        for (int j = 0; j < PV_MAX_COLS; ++j) v = (j < p.n_cols && p.col_target[j] == i) ? pc[j] : v;
                        th[i] = v;
                    }
                    double y0[M::NX];
                    M::hoist(th, ks.c, y0, sink.obs);
                    bool over[M::NX];
#pragma unroll
                    for (int i = 0; i < M::NX; ++i) {
                        double v = sh_y0[i];
                        over[i] = (v == v);
                        v = over[i] ? v : y0[i];
#pragma unroll
                        for (int j = 0; j < PV_MAX_COLS; ++j)
                            if (j < p.n_cols && p.col_target[j] == M::NTH + i) {
                                v = pc[j];
                                over[i] = true;
                            }
                        z[i] = v;
                    }
                    if constexpr (SENS) {
                        double dy0[M::NSD][M::NX];
                        M::hoist_sens(th, ks.dc, dy0);
#pragma unroll
                        for (int s = 0; s < M::NS; ++s)
#pragma unroll
                            for (int i = 0; i < M::NX; ++i) z[M::NX * (1 + s) + i] = over[i] ? 0.0 : dy0[s][i];
                    }
                    sink.begin(b, qbase);
                    have = true;
                    if constexpr (METHOD == 1) lane.stiff_after = p.stiff_after;
                    const int rc = lane.start(ks, z, p.t0, sh_t, p.T, p.rtol, p.atol, sink);
                    if (rc != PV_RUNNING) finish(rc);
                }
            }
        }
        if (__all_sync(FULL, !have)) {
            if (__all_sync(FULL, exhausted)) break;
            continue;
        }
        if (have) {
            const int rc = lane.step(ks, sh_t, p.T, p.rtol, p.atol, p.max_steps, sink);
            if (rc != PV_RUNNING) finish(rc);
        }
    }
}

// ---- the lockstep kernel -------------------------------------------------------------------------------------
// One warp = 32 CONSECUTIVE work items (neighbours in parameter space when the launch carries the locality schedule's
// index map) integrated with one common step size (pv_dopri5_lane::step_lockstep).  Everything that makes
// pv_batch_kernel's warps diverge is warp-uniform here: t, h, accept / reject, the grid crossings, start and end of a
// trajectory - so there is no refill machinery, the per-grid-point code runs with all 32 lanes, and a warp executes
// (steps of its most demanding lane) instead of (sum over iterations of whichever lanes are busy).  A lane's result now
// depends on its 31 neighbours' step sizes - within the tolerances, but not bit-identical across batch compositions; the
// work-item order of a launch is deterministic (pv_schedule.cuh sorts every bucket), so a given batch reproduces itself.
#ifndef PV_MINB_LOCKSTEP
#define PV_MINB_LOCKSTEP PV_MINB_SMALL
#endif
// trajectory stores go through the per-warp staging block (pv_sink STAGE) when there is no sensitivity output
template <int MODE>
constexpr bool pv_lockstep_stage = PV_STAGE_ON && (MODE & PV_MODE_TRAJ) != 0 && (MODE & PV_MODE_SENS) == 0;
template <int MODE>
constexpr size_t pv_lockstep_stage_bytes = pv_lockstep_stage<MODE> ? (size_t)(PV_BLOCK / 32) * 32 * PV_STAGE_RS * sizeof(double) : 0;
template <class M, int MODE, bool OBSID, bool QSH>
__global__ void __launch_bounds__(PV_BLOCK, (pv_system<M, (MODE & PV_MODE_SENS) != 0>::N <= 4) ? PV_MINB_LOCKSTEP : 1)
pv_batch_lockstep_kernel(const __grid_constant__ pv_launch_params_m<M> p) {
    constexpr bool SENS = (MODE & PV_MODE_SENS) != 0;
    constexpr bool LOGP = (MODE & PV_MODE_LOGP) != 0;
    using Sys = pv_system<M, SENS>;
    using Lane = pv_dopri5_warp<Sys>;
    extern __shared__ __align__(16) double sh[];
    constexpr bool stats_shared = LOGP && QSH;     // QSH <=> n_data == 1 (the host dispatches on it)
    constexpr bool STAGE = pv_lockstep_stage<MODE>;
    double* sh_stats = sh;
    double* sh_t = sh + (stats_shared ? PV_QSTRIDE * p.T * p.n_obs : 0);
    double* sh_theta = sh_t + p.T;
    double* sh_y0 = sh_theta + M::NTH;
    double* sh_stage = sh_y0 + M::NX;              // [warps][32][PV_STAGE_RS] when STAGE (pv_lockstep_stage_bytes)
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) sh_t[i] = p.tgrid[i];
    for (int i = threadIdx.x; i < M::NTH; i += blockDim.x) sh_theta[i] = p.theta_base[i];
    for (int i = threadIdx.x; i < M::NX; i += blockDim.x) sh_y0[i] = p.y0_base[i];
    if (stats_shared)
        for (int i = threadIdx.x; i < PV_QSTRIDE * p.T * p.n_obs; i += blockDim.x) sh_stats[i] = p.stats[i];
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    const int64_t n_items = p.n_items_dev ? (int64_t)*p.n_items_dev : p.B;
    const double* qbase = stats_shared ? sh_stats : p.stats;
    typename Sys::Consts ks;
    Lane lane;
    lane.stiff_after = p.stiff_after;
    pv_sink<M, MODE, OBSID, stats_shared, STAGE> sink{p, 0, nullptr, nullptr, nullptr};
    if constexpr (STAGE) sink.st_base = (unsigned)__cvta_generic_to_shared(sh_stage + (threadIdx.x >> 5) * (32 * PV_STAGE_RS));
    for (;;) {
        unsigned long long base = 0;
        if (lane_id == 0) base = atomicAdd(p.counter, 32ull);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= n_items) break;
        // Tiles are handed out from the END of the work list: under the locality schedule that is the high-key end (largest
        // rate constants = most steps), so the longest tiles start first and the launch does not end on a 1700-step tile
        // that was picked up last (longest-processing-time-first).
        // lanes past the end of the work list integrate a copy of the last item (identical values to identical addresses)
        const int64_t tile0 = (n_items - 1) / 32 * 32 - (int64_t)base;
        const int64_t item = min(tile0 + lane_id, n_items - 1);
        const int64_t b = p.index_map ? (int64_t)p.index_map[item] : item;
        double pc[PV_MAX_COLS];
#pragma unroll
        for (int j = 0; j < PV_MAX_COLS; ++j) pc[j] = (j < p.n_cols) ? __ldg(p.P + (int64_t)j * p.B + b) : 0.0;
        double z[Sys::N];
        double th[M::NTH];
#pragma unroll
        for (int i = 0; i < M::NTH; ++i) {
            double v = sh_theta[i];
#pragma unroll
            for (int j = 0; j < PV_MAX_COLS; ++j) v = (j < p.n_cols && p.col_target[j] == i) ? pc[j] : v;
            th[i] = v;
        }
        double y0[M::NX];
        M::hoist(th, ks.c, y0, sink.obs);
        bool over[M::NX];
#pragma unroll
        for (int i = 0; i < M::NX; ++i) {
            double v = sh_y0[i];
            over[i] = (v == v);
            v = over[i] ? v : y0[i];
#pragma unroll
            for (int j = 0; j < PV_MAX_COLS; ++j)
                if (j < p.n_cols && p.col_target[j] == M::NTH + i) {
                    v = pc[j];
                    over[i] = true;
                }
            z[i] = v;
        }
        if constexpr (SENS) {
            double dy0[M::NSD][M::NX];
            M::hoist_sens(th, ks.dc, dy0);
#pragma unroll
            for (int s = 0; s < M::NS; ++s)
#pragma unroll
                for (int i = 0; i < M::NX; ++i) z[M::NX * (1 + s) + i] = over[i] ? 0.0 : dy0[s][i];
        }
        sink.begin(b, qbase);
        int rc = lane.start(ks, z, p.t0, sh_t, p.T, p.rtol, p.atol, sink);     // uniform: same t0 and grid for every lane
        while (rc == PV_RUNNING)           // an accepted step flips lane.cur: a warp-uniform jump to the other copy of the step
            rc = lane.cur ? lane.template step<1>(ks, sh_t, p.T, p.rtol, p.atol, p.max_steps, sink)
                          : lane.template step<0>(ks, sh_t, p.T, p.rtol, p.atol, p.max_steps, sink);
        sink.flush();                      // the grid points still staged (STAGE)
        if (lane.dead) rc = PV_ERR_NONFINITE;
        if constexpr (LOGP) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            p.logp[b] = (rc == 0) ? sink.lp : nan;
            if constexpr (SENS) {
#pragma unroll
                for (int s = 0; s < M::NS; ++s) p.grad[(int64_t)s * p.B + b] = (rc == 0) ? sink.g[s] : nan;
                if constexpr ((MODE & PV_MODE_FIM) != 0) {
#pragma unroll
                    for (int q = 0; q < M::NS * (M::NS + 1) / 2; ++q) p.fim[(int64_t)q * p.B + b] = (rc == 0) ? sink.fim[q] : nan;
                }
            }
        }
        if (p.status) p.status[b] = rc;
        if (p.steps) {
            p.steps[b] = lane.st.accepted;
            p.steps[p.B + b] = lane.st.rejected;
        }
    }
}

#include "pv_dopri5_pairs.cuh"

// ---- deterministic segment sum: one warp per segment, fixed association order --------------------------------
__global__ void pv_segment_sum_kernel(const double* __restrict__ x, int64_t seg_len, int64_t n_seg,
                                      double* __restrict__ out) {
    const int64_t seg = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (seg >= n_seg) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int64_t i = lane; i < seg_len; i += 32) acc += x[seg * seg_len + i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[seg] = acc;
}

// work list of the trajectories whose status code is in code_mask (bit k = code k) (order irrelevant: results go to their own slots)
__global__ void pv_compact_status_kernel(const int32_t* __restrict__ status, int64_t B, unsigned code_mask, int32_t* __restrict__ index,
                                         int* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = i < B && ((code_mask >> (status[i] & 31)) & 1u);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (hit) index[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
}

// count trajectories with status != 0 (for the *_host return value)
__global__ void pv_count_failed_kernel(const int32_t* __restrict__ status, int64_t B, int* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = (i < B && status[i] != 0) ? 1 : 0;
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(out, bad);
}

// ---- FMA-pipe peak microbenchmark (roofline denominator) --------------------------------
template <class Tp>
__global__ void pv_fma_peak_kernel(Tp* out, int iters, Tp a, Tp bq) {
    Tp x0 = threadIdx.x * (Tp)1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
       x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, bq); x1 = fma(x1, a, bq); x2 = fma(x2, a, bq); x3 = fma(x3, a, bq);
            x4 = fma(x4, a, bq); x5 = fma(x5, a, bq); x6 = fma(x6, a, bq); x7 = fma(x7, a, bq);
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
