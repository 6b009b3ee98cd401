// C ABI of libparvar_b200.so (include/parvar_b200.h).  Host-side plumbing only: argument checks,
// device-side copies of the small configuration arrays, kernel dispatch over
// (model x integrator x mode), chunked host<->device pipelines.  No arithmetic of the path runs
// on the host, and there is no CPU fallback: without a device every compute call returns PV_E_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/parvar_b200.h"
#include "pv_sampler_host.cuh"
#include "pv_kernels.cuh"
#include "pv_rosenbrock_split.cuh"
#include "pv_hier.cuh"
#include "pv_hmc.cuh"
#include "pv_nuts.cuh"
#include "pv_schedule.cuh"
#include "pv_sampler_device.cuh"
#include "pv_rng_host.cuh"
#include "pv_optimizer_device.cuh"
// the models compiled into this library: the three registry models by default; a run-time build of another SBML file
// (codegen.compile_model: generate -> nvcc -> cached .so) points PV_MODELS_INC at its own list
#ifndef PV_MODELS_INC
#define PV_MODELS_INC "generated/models.inc"
#endif
#include PV_MODELS_INC

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_OK(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(PV_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    } while (0)

struct pv_problem {
    int model = -1;
    int device = 0;
    const pv_model_table* tab = nullptr;
    int method = PV_METHOD_AUTO;
    double rtol = 1e-7, atol = 1e-10;
    int max_steps = 1000000;
    std::vector<double> theta_base, y0_base, tgrid, stats, sigma;
    std::vector<int> col_target, obs_idx;
    int noise_model = PV_NOISE_NORMAL;
    int n_data = 0;
    // device copies
    double *d_tgrid = nullptr, *d_stats = nullptr;   // (batch-wide values travel in the kernel parameters)
    // host pipeline
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;                               // all device->host copies of the host pipeline, FIFO
    cudaEvent_t ev_ready[3] = {nullptr, nullptr, nullptr};            // slot's kernels done
    cudaEvent_t ev_drained[3] = {nullptr, nullptr, nullptr};          // slot's results copied out
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double* stage_dev[3] = {nullptr, nullptr, nullptr};
    size_t stage_bytes[3] = {0, 0, 0};
    int* d_failed = nullptr;
    double last_kernel_ms = 0.0;
    // lane-refill work counters: one per launch in flight (ring), zeroed in-stream before each launch
    unsigned long long* d_counters = nullptr;
    int counter_next = 0;
    unsigned long long* own_counters = nullptr;   // set while a sampler captures its iteration graph: that graph's private slots
    int own_counter_next = 0;
    int num_sms = 0;
    int refill_min = 0;   // 0 = per-integrator default (DOPRI5: 6, RODAS4: 32), see pv_problem_set_schedule
    int layout = PV_LAYOUT_AUTO;   // see pv_problem_set_layout
    int max_ctas_per_sm = 0;       // 0 = what the occupancy calculator allows (pv_problem_set_tuning)
    int64_t host_chunk_rows = (int64_t)1 << 15;
    double* d_hier = nullptr;      // logp [B] + grad [NS][B] scratch of pv_hier_logp_grad
    size_t hier_bytes = 0;
    double t0 = 0.0;
    // METHOD_AUTO on a non-stiff model: trajectories that exhaust switch_steps explicit steps (stiff parameter
    // values) are re-run with RODAS4.  Scratch per launch slot: 0 = caller's stream, 1..3 = internal pipeline streams.
    int switch_steps = 4000;
    int locality = 0;              // pv_problem_set_locality: 0 = auto (sorted launches from 32 768 trajectories), 1 = off, 2 = on
    int32_t* d_sched_index[4] = {nullptr, nullptr, nullptr, nullptr};     // locality-schedule scratch per launch slot (pv_schedule.cuh)
    int lockstep = 0;              // pv_problem_set_lockstep: 0 = auto, 1 = off, 2 = on
    int roles = 0;                 // pv_problem_set_group_lanes: 0 = auto, 1 = off, 2 = on
    int64_t sched_cap[4] = {0, 0, 0, 0};
    int stiff_after = 128;         // DOPRI5 stiffness detection of the AUTO first pass starts after this many steps
    int32_t* d_scratch_status[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t* d_scratch_index[4] = {nullptr, nullptr, nullptr, nullptr};
    int* d_scratch_count[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t scratch_cap[4] = {0, 0, 0, 0};
};
#define PV_COUNTER_RING 1024
#define PV_SAMPLER_COUNTERS 8
#ifndef PV_LOCKSTEP_AUTO
#define PV_LOCKSTEP_AUTO 1     // measured default of pv_problem_set_lockstep(0) for sorted launches (tools/bench_configs.py)
#endif

// ---------------------------------------------------------------------------------------------------------
extern "C" int pv_abi_version(void) { return PV_ABI_VERSION; }
extern "C" const char* pv_last_error(void) { return g_last_error.c_str(); }
extern "C" int pv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" int pv_model_count(void) { return PV_N_MODELS; }
extern "C" const char* pv_model_name(int i) { return (i >= 0 && i < PV_N_MODELS) ? pv_model_tables[i].name : nullptr; }
extern "C" int64_t pv_launch_count(void) { return g_launches.load(); }

static int find_id(const char* const* ids, int n, const std::string& id) {
    for (int i = 0; i < n; ++i)
        if (id == ids[i]) return i;
    return -1;
}
static std::string strip_brackets(const char* id) {
    std::string s(id ? id : "");
    if (s.size() >= 2 && s.front() == '[' && s.back() == ']') s = s.substr(1, s.size() - 2);
    return s;
}

extern "C" pv_problem* pv_problem_create(const char* model_name, int device) {
    if (!model_name) {
        fail(PV_E_ARG, "model_name is NULL");
        return nullptr;
    }
    int m = -1;
    for (int i = 0; i < PV_N_MODELS; ++i)
        if (!strcmp(pv_model_tables[i].name, model_name)) m = i;
    if (m < 0) {
        fail(PV_E_MODEL, std::string("unknown model '") + model_name + "' (not compiled into this library)");
        return nullptr;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        fail(PV_E_CUDA, "no CUDA device: parvar_b200 has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= n) {
        fail(PV_E_ARG, "device index out of range");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(PV_E_CUDA, "cudaSetDevice failed");
        return nullptr;
    }
    auto* p = new pv_problem();
    p->model = m;
    p->device = device;
    p->tab = &pv_model_tables[m];
    p->theta_base.assign(p->tab->theta_default, p->tab->theta_default + p->tab->n_theta);
    p->y0_base.assign(p->tab->n_states, std::numeric_limits<double>::quiet_NaN());
    bool ok = cudaMalloc(&p->d_failed, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&p->d_counters, sizeof(unsigned long long) * PV_COUNTER_RING) == cudaSuccess &&
              cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    for (int i = 0; i < 3 && ok; ++i)
        ok = cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->ev_ready[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->ev_drained[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&p->ev0) == cudaSuccess && cudaEventCreate(&p->ev1) == cudaSuccess;
    if (!ok) {
        fail(PV_E_CUDA, std::string("problem setup: ") + cudaGetErrorString(cudaGetLastError()));
        pv_problem_destroy(p);
        return nullptr;
    }
    return p;
}

extern "C" void pv_problem_destroy(pv_problem* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_hier);
    cudaFree(p->d_tgrid);
    cudaFree(p->d_stats);
    cudaFree(p->d_failed);
    cudaFree(p->d_counters);
    for (int i = 0; i < 4; ++i) {
        cudaFree(p->d_scratch_status[i]);
        cudaFree(p->d_scratch_index[i]);
        cudaFree(p->d_scratch_count[i]);
        cudaFree(p->d_sched_index[i]);
    }
    for (int i = 0; i < 3; ++i) {
        if (p->streams[i]) cudaStreamDestroy(p->streams[i]);
        if (p->ev_ready[i]) cudaEventDestroy(p->ev_ready[i]);
        if (p->ev_drained[i]) cudaEventDestroy(p->ev_drained[i]);
        cudaFree(p->stage_dev[i]);
    }
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}

extern "C" int pv_problem_dims(const pv_problem* p, int* nx, int* nth, int* ns) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (nx) *nx = p->tab->n_states;
    if (nth) *nth = p->tab->n_theta;
    if (ns) *ns = p->tab->n_sens;
    return 0;
}
extern "C" const char* pv_problem_state_id(const pv_problem* p, int i) {
    return (p && i >= 0 && i < p->tab->n_states) ? p->tab->state_ids[i] : nullptr;
}
extern "C" const char* pv_problem_theta_id(const pv_problem* p, int i) {
    return (p && i >= 0 && i < p->tab->n_theta) ? p->tab->theta_ids[i] : nullptr;
}
extern "C" double pv_problem_theta_default(const pv_problem* p, int i) {
    return (p && i >= 0 && i < p->tab->n_theta) ? p->tab->theta_default[i] : std::numeric_limits<double>::quiet_NaN();
}
extern "C" int pv_problem_sens_theta_index(const pv_problem* p, int k) {
    return (p && k >= 0 && k < p->tab->n_sens) ? p->tab->sens_index[k] : PV_E_ARG;
}
extern "C" int pv_problem_flops(const pv_problem* p, const char* what) {
    if (!p || !what) return fail(PV_E_ARG, "NULL argument");
    for (int i = 0; i < p->tab->n_flops; ++i)
        if (!strcmp(p->tab->flop_names[i], what)) return p->tab->flop_values[i];
    return fail(PV_E_ARG, std::string("unknown flop counter '") + what + "'");
}

extern "C" int pv_problem_set_schedule(pv_problem* p, int refill_min) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (refill_min < 0 || refill_min > 32) return fail(PV_E_ARG, "refill_min must be in [0, 32] (0 = default)");
    p->refill_min = refill_min;
    return 0;
}

extern "C" int pv_problem_set_layout(pv_problem* p, int layout) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (layout < PV_LAYOUT_AUTO || layout > PV_LAYOUT_SPLIT) return fail(PV_E_ARG, "unknown layout");
    p->layout = layout;
    return 0;
}

extern "C" int pv_problem_set_tuning(pv_problem* p, int max_ctas_per_sm, int64_t host_chunk_rows) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (max_ctas_per_sm < 0) return fail(PV_E_ARG, "max_ctas_per_sm must be >= 0 (0 = no cap)");
    if (host_chunk_rows != 0 && host_chunk_rows < 1024) return fail(PV_E_ARG, "host_chunk_rows must be 0 (default) or >= 1024");
    p->max_ctas_per_sm = max_ctas_per_sm;
    p->host_chunk_rows = host_chunk_rows ? host_chunk_rows : ((int64_t)1 << 15);
    return 0;
}

extern "C" int pv_problem_set_stiff_switch(pv_problem* p, int switch_steps) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (switch_steps < 1) return fail(PV_E_ARG, "switch_steps must be >= 1");
    p->switch_steps = switch_steps;
    return 0;
}

extern "C" int pv_problem_set_locality(pv_problem* p, int mode) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (mode < 0 || mode > 2) return fail(PV_E_ARG, "mode must be 0 (auto), 1 (off) or 2 (on)");
    p->locality = mode;
    return 0;
}

extern "C" int pv_problem_set_lockstep(pv_problem* p, int mode) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (mode < 0 || mode > 2) return fail(PV_E_ARG, "mode must be 0 (auto), 1 (off) or 2 (on)");
    p->lockstep = mode;
    return 0;
}

extern "C" int pv_problem_set_group_lanes(pv_problem* p, int mode) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (mode < 0 || mode > 2) return fail(PV_E_ARG, "mode must be 0 (auto), 1 (off) or 2 (on)");
    p->roles = mode;
    return 0;
}

extern "C" int pv_problem_set_stiffness_detection(pv_problem* p, int after_steps) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (after_steps < 0) return fail(PV_E_ARG, "after_steps must be >= 0");
    p->stiff_after = after_steps;
    return 0;
}

extern "C" int pv_problem_set_solver(pv_problem* p, int method, double rtol, double atol, int max_steps) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (method < PV_METHOD_AUTO || method > PV_METHOD_RODAS4) return fail(PV_E_ARG, "unknown method");
    if (!(rtol > 0) || !(atol >= 0) || max_steps <= 0) return fail(PV_E_ARG, "rtol > 0, atol >= 0, max_steps > 0 required");
    p->method = method;
    p->rtol = rtol;
    p->atol = atol;
    p->max_steps = max_steps;
    return 0;
}

// returns theta index, or n_theta + state index, or -1
static int resolve_id(const pv_problem* p, const char* id) {
    const std::string s = strip_brackets(id);
    int i = find_id(p->tab->theta_ids, p->tab->n_theta, s);
    if (i >= 0) return i;
    i = find_id(p->tab->state_ids, p->tab->n_states, s);
    if (i >= 0) return p->tab->n_theta + i;
    return -1;
}

extern "C" int pv_problem_set_value(pv_problem* p, const char* id, double value) {
    if (!p || !id) return fail(PV_E_ARG, "NULL argument");
    const int t = resolve_id(p, id);
    if (t < 0) return fail(PV_E_MODEL, std::string("model '") + p->tab->name + "' has no settable id '" + id + "'");
    if (t < p->tab->n_theta)
        p->theta_base[t] = value;
    else
        p->y0_base[t - p->tab->n_theta] = value;
    return 0;
}
extern "C" int pv_problem_reset_values(pv_problem* p) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    p->theta_base.assign(p->tab->theta_default, p->tab->theta_default + p->tab->n_theta);
    p->y0_base.assign(p->tab->n_states, std::numeric_limits<double>::quiet_NaN());
    return 0;
}

extern "C" int pv_problem_set_columns(pv_problem* p, int n_cols, const char* const* ids) {
    if (!p || n_cols < 0 || (n_cols > 0 && !ids)) return fail(PV_E_ARG, "bad columns");
    if (n_cols > PV_MAX_COLS) return fail(PV_E_ARG, "too many per-trajectory columns (max 8)");
    std::vector<int> tgt;
    for (int j = 0; j < n_cols; ++j) {
        const int t = resolve_id(p, ids[j]);
        if (t < 0) return fail(PV_E_MODEL, std::string("model '") + p->tab->name + "' has no settable id '" + ids[j] + "'");
        tgt.push_back(t);
    }
    p->col_target = tgt;
    return 0;
}

extern "C" int pv_problem_set_timepoints(pv_problem* p, double t0, int n_t, const double* t_host) {
    if (!p || n_t <= 0 || !t_host) return fail(PV_E_ARG, "bad timepoints");
    if (!(t0 <= t_host[0])) return fail(PV_E_ARG, "t0 must be <= the first timepoint");
    for (int i = 1; i < n_t; ++i)
        if (!(t_host[i] > t_host[i - 1])) return fail(PV_E_ARG, "timepoints must be strictly increasing");
    CUDA_OK(cudaSetDevice(p->device));
    p->tgrid.assign(t_host, t_host + n_t);
    p->t0 = t0;
    cudaFree(p->d_tgrid);
    p->d_tgrid = nullptr;
    CUDA_OK(cudaMalloc(&p->d_tgrid, sizeof(double) * n_t));
    CUDA_OK(cudaMemcpy(p->d_tgrid, t_host, sizeof(double) * n_t, cudaMemcpyHostToDevice));
    // a pageable-source copy may return once the data are staged: wait for the DMA itself, because the kernels run on
    // non-blocking streams that do not order against the legacy stream (configuration time, not the hot loop)
    CUDA_OK(cudaStreamSynchronize(0));
    p->stats.clear();   // data are tied to the grid
    p->n_data = 0;
    return 0;
}

extern "C" int pv_problem_set_observables(pv_problem* p, int n_obs, const char* const* ids) {
    if (!p || n_obs <= 0 || !ids) return fail(PV_E_ARG, "bad observables");
    if (n_obs > PV_MAX_OBS) return fail(PV_E_ARG, "too many observables (max 16)");
    std::vector<int> idx;
    for (int k = 0; k < n_obs; ++k) {
        const int i = find_id(p->tab->state_ids, p->tab->n_states, strip_brackets(ids[k]));
        if (i < 0) return fail(PV_E_MODEL, std::string("model '") + p->tab->name + "' has no state '" + ids[k] + "'");
        idx.push_back(i);
    }
    p->obs_idx = idx;
    p->stats.clear();
    p->n_data = 0;
    return 0;
}

extern "C" int pv_problem_set_data(pv_problem* p, int noise_model, int n_data, const double* stats_host,
                                   const double* sigma_host) {
    if (!p || !stats_host || !sigma_host || n_data <= 0) return fail(PV_E_ARG, "bad data");
    if (noise_model != PV_NOISE_NORMAL && noise_model != PV_NOISE_LOGNORMAL) return fail(PV_E_ARG, "unknown noise model");
    if (p->tgrid.empty() || p->obs_idx.empty()) return fail(PV_E_STATE, "set timepoints and observables before data");
    const size_t n = (size_t)3 * p->tgrid.size() * p->obs_idx.size() * (size_t)n_data;
    for (size_t k = 0; k < p->obs_idx.size(); ++k)
        if (!(sigma_host[k] > 0)) return fail(PV_E_ARG, "sigma must be > 0");
    CUDA_OK(cudaSetDevice(p->device));
    p->stats.assign(stats_host, stats_host + n);
    p->sigma.assign(sigma_host, sigma_host + p->obs_idx.size());
    p->noise_model = noise_model;
    p->n_data = n_data;
    // The kernels do not see (n, S1, S2) but the coefficients of the likelihood term as a polynomial in the model
    // output v (v = f for normal noise, v = log f for log-normal noise):
    //   lp(j,k) = q0 + v (q1 + q2 v),   q2 = -n/(2 sigma^2),  q1 = S1/sigma^2,
    //   q0 = -(n/2) log(2 pi sigma^2) - S2/(2 sigma^2)            (normal)
    //   q0 = -S1 - n (log sigma + log(2 pi)/2) - S2/(2 sigma^2)   (log-normal; -S1 = -sum log y)
    // so an entry without measurements (n = 0) contributes exactly 0 without a branch; d lp / d v = q1 + 2 q2 v and the
    // Fisher weight n / sigma^2 = -2 q2.
    // Device layout: one record (q0, q1, q2, pad) per (data set, time, observable), data-set-major [n_data][T][n_obs][4]:
    // a lane reads the records of its trajectory's data set front to back (two 16-byte loads each) as it crosses the grid.
    const double LOG_2PI = 1.8378770664093454835606594728112;
    const size_t T = p->tgrid.size(), K = p->obs_idx.size(), D = (size_t)n_data, plane = T * K * D;
    const size_t nq = PV_QSTRIDE * plane;
    std::vector<double> q(nq, 0.0);
    for (size_t j = 0; j < T; ++j)
        for (size_t k = 0; k < K; ++k) {
            const double sg = sigma_host[k], inv = 1.0 / (sg * sg);
            for (size_t d = 0; d < D; ++d) {
                const size_t o = (j * K + k) * D + d;
                const double nn = stats_host[o], s1 = stats_host[plane + o], s2 = stats_host[2 * plane + o];
                double* r = q.data() + ((d * T + j) * K + k) * PV_QSTRIDE;
                r[2] = -0.5 * nn * inv;
                r[1] = s1 * inv;
                r[0] = (noise_model == PV_NOISE_NORMAL ? -0.5 * nn * (LOG_2PI + 2.0 * std::log(sg))
                                                       : -s1 - nn * (std::log(sg) + 0.5 * LOG_2PI)) - 0.5 * s2 * inv;
            }
        }
    cudaFree(p->d_stats);
    p->d_stats = nullptr;
    CUDA_OK(cudaMalloc(&p->d_stats, sizeof(double) * nq));
    CUDA_OK(cudaMemcpy(p->d_stats, q.data(), sizeof(double) * nq, cudaMemcpyHostToDevice));
    CUDA_OK(cudaStreamSynchronize(0));   // see pv_problem_set_timepoints
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------------------
constexpr int PV_MAX_DYN_SMEM = 200 * 1024;   // pv_launch refuses time grids / data that need more

// resident CTAs per SM of (kernel, block size, dynamic shared memory): asked once, then cached - the occupancy query
// and the opt-in attribute cost several microseconds each, which the small latency-bound launches of the sampler
// loops would pay at every iteration
static int resident_ctas(const void* kern, int threads, size_t shmem, cudaError_t& err) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, size_t, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(kern, threads, shmem, dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        err = cudaSuccess;
        return it->second;
    }
    // The opt-in limit is a property of the FUNCTION, shared by every problem and host thread that launches it: it is set to
    // one constant (the largest size pv_launch accepts), never to this launch's size - otherwise a thread with a short
    // time grid could lower it between another thread's cudaFuncSetAttribute and its launch ("invalid argument").
    if (shmem > 48 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PV_MAX_DYN_SMEM);
        if (err != cudaSuccess) return 0;
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, shmem);
    if (err != cudaSuccess) return 0;
    if (per_sm < 1) {
        err = cudaErrorLaunchOutOfResources;
        return 0;
    }
    cache[key] = per_sm;
    return per_sm;
}

struct pv_launch_ctx {
    int num_sms, layout, max_ctas_per_sm, lockstep, roles;
    const double* theta;   // host [NTH]
    const double* y0;      // host [NX]
};

template <class M>
static pv_launch_params_m<M> with_values(const pv_launch_params& lp, const pv_launch_ctx& cx) {
    pv_launch_params_m<M> out;
    static_cast<pv_launch_params&>(out) = lp;
    for (int i = 0; i < M::NTH; ++i) out.theta_base[i] = cx.theta[i];
    for (int i = 0; i < M::NX; ++i) out.y0_base[i] = cx.y0[i];
    return out;
}

template <class M, class KernelT>
static cudaError_t launch_persistent(KernelT kern, const pv_launch_params& lp, const pv_launch_ctx& cx, int threads,
                                     int64_t items_per_cta, size_t shmem, cudaStream_t s) {
    // persistent grid: as many CTAs as are resident at once (a multiple of the SM count), never more than the work
    cudaError_t e = cudaSuccess;
    int per_sm = resident_ctas((const void*)kern, threads, shmem, e);
    if (e != cudaSuccess) return e;
    if (cx.max_ctas_per_sm >= 1 && cx.max_ctas_per_sm < per_sm) per_sm = cx.max_ctas_per_sm;
    const int64_t want = (lp.B + items_per_cta - 1) / items_per_cta;
    const int64_t blocks = std::min<int64_t>(want, (int64_t)cx.num_sms * per_sm);
    kern<<<(unsigned)blocks, threads, shmem, s>>>(with_values<M>(lp, cx));
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <class M, int METHOD, int MODE, bool OBSID>
static cudaError_t launch_one(const pv_launch_params& lp, const pv_launch_ctx& cx, size_t shmem, cudaStream_t s) {
    if constexpr (METHOD == 1) {
        // (a time grid / data set so large that the staging block no longer fits keeps the per-lane kernel)
        if (cx.lockstep && shmem + pv_lockstep_stage_bytes<MODE> <= (size_t)PV_MAX_DYN_SMEM) {
            shmem += pv_lockstep_stage_bytes<MODE>;
            if ((MODE & PV_MODE_LOGP) && lp.n_data == 1)
                return launch_persistent<M>(pv_batch_lockstep_kernel<M, MODE, OBSID, (MODE & PV_MODE_LOGP) != 0>, lp, cx, PV_BLOCK, PV_BLOCK, shmem, s);
            return launch_persistent<M>(pv_batch_lockstep_kernel<M, MODE, OBSID, false>, lp, cx, PV_BLOCK, PV_BLOCK, shmem, s);
        }
    }
    return launch_persistent<M>(pv_batch_kernel<M, METHOD, MODE, OBSID>, lp, cx, PV_BLOCK, PV_BLOCK, shmem, s);
}

// RODAS4 + sensitivities of the large models: one warp per block of the augmented system (pv_rosenbrock_split.cuh)
template <class M>
constexpr bool pv_use_split = M::NS >= 1 && M::NS <= 2 && M::NX * (1 + M::NS) > 16;

template <class M>
static cudaError_t launch_model(int method, int mode, bool obsid, const pv_launch_params& lp, const pv_launch_ctx& cx,
                                size_t shmem, cudaStream_t s) {
    if constexpr (pv_use_split<M>) {
        // measured (tools/bench_configs.py, icg_body_flat logp + gradient, stage vectors in shared memory): 8000
        // trajectories 3.6 ms split vs 11 ms lane; 128 000 trajectories 49 ms split vs 68 ms lane -> AUTO = split
        if (method == 2 && (mode == PV_MODE_LOGP_GRAD || mode == PV_MODE_LOGP_GRAD_FIM) && cx.layout != PV_LAYOUT_LANE &&
            pv_split_layout<M>::bytes(lp.T, lp.n_obs) <= 100 * 1024) {
            const size_t sh = pv_split_layout<M>::bytes(lp.T, lp.n_obs);
            const int threads = 32 * (1 + M::NS);
            return mode == PV_MODE_LOGP_GRAD
                       ? launch_persistent<M>(pv_rodas4_split_kernel<M, PV_MODE_LOGP_GRAD>, lp, cx, threads, 32, sh, s)
                       : launch_persistent<M>(pv_rodas4_split_kernel<M, PV_MODE_LOGP_GRAD_FIM>, lp, cx, threads, 32, sh, s);
        }
    }
    // small batches of DOPRI5 logp + gradient: one trajectory per GROUP of NS lanes (pv_dopri5_pairs.cuh) while the
    // launch fits one wave of that kernel - beyond it the lane-per-trajectory kernels' throughput wins
    if constexpr (M::NS >= 1) {
        // one trajectory per group of NS lanes, one-warp CTAs (pv_dopri5_pairs.cuh); "one wave" = 4 warps per scheduler
        constexpr int TPW = 32 / M::NS;
        const int64_t one_wave = (int64_t)cx.num_sms * 16 * TPW;
        if (method == 1 && mode == PV_MODE_LOGP_GRAD && (cx.roles == 2 || (cx.roles == 0 && lp.B <= one_wave))) {
            if (lp.n_data == 1) return launch_persistent<M>(pv_batch_pairs_kernel<M, true>, lp, cx, PV_PAIRS_BLOCK, TPW, shmem, s);
            return launch_persistent<M>(pv_batch_pairs_kernel<M, false>, lp, cx, PV_PAIRS_BLOCK, TPW, shmem, s);
        }
    }
    // the all-states-in-order specialisation is compiled for the explicit integrator only (cheap steps, where the
    // per-output code matters); the Rosenbrock kernels always use the generic sink
#define PV_CASE(METH, MODE)                                                                \
    if (method == METH && mode == MODE) {                                                  \
        if (METH == 1 && obsid) return launch_one<M, METH, MODE, (METH == 1)>(lp, cx, shmem, s); \
        return launch_one<M, METH, MODE, false>(lp, cx, shmem, s);                         \
    }
    PV_CASE(1, PV_MODE_TRAJ) PV_CASE(1, PV_MODE_LOGP) PV_CASE(1, PV_MODE_TRAJ_LOGP) PV_CASE(1, PV_MODE_LOGP_GRAD) PV_CASE(1, PV_MODE_TRAJ_SENS)
    PV_CASE(1, PV_MODE_LOGP_GRAD_FIM)
    PV_CASE(2, PV_MODE_TRAJ) PV_CASE(2, PV_MODE_LOGP) PV_CASE(2, PV_MODE_TRAJ_LOGP) PV_CASE(2, PV_MODE_LOGP_GRAD) PV_CASE(2, PV_MODE_TRAJ_SENS)
    PV_CASE(2, PV_MODE_LOGP_GRAD_FIM)
#undef PV_CASE
    return cudaErrorInvalidValue;
}

static int ensure_scratch(pv_problem* p, int slot, int64_t B) {
    if (p->scratch_cap[slot] >= B) return 0;
    cudaFree(p->d_scratch_status[slot]);
    cudaFree(p->d_scratch_index[slot]);
    p->d_scratch_status[slot] = p->d_scratch_index[slot] = nullptr;
    p->scratch_cap[slot] = 0;
    CUDA_OK(cudaMalloc(&p->d_scratch_status[slot], sizeof(int32_t) * B));
    CUDA_OK(cudaMalloc(&p->d_scratch_index[slot], sizeof(int32_t) * B));
    if (!p->d_scratch_count[slot]) CUDA_OK(cudaMalloc(&p->d_scratch_count[slot], sizeof(int)));
    p->scratch_cap[slot] = B;
    return 0;
}

// lane-refill work counter of the next launch: a slot of the problem's ring, or - while a sampler records its iteration
// graph, whose replays must not share a slot with launches made meanwhile - one of the graph's own slots
static unsigned long long* next_counter(pv_problem* p) {
    if (p->own_counters) return p->own_counters + (p->own_counter_next++ % PV_SAMPLER_COUNTERS);
    unsigned long long* c = p->d_counters + p->counter_next;
    p->counter_next = (p->counter_next + 1) % PV_COUNTER_RING;
    return c;
}

// locality schedule of one launch (pv_schedule.cuh): fills the slot's index map on stream s
static int locality_schedule(pv_problem* p, int slot, int64_t B, const double* P, cudaStream_t s, const int32_t** index_out) {
    const size_t per = ((size_t)B + 63) / 64 * 64;    // elements per array, 256-byte multiples
    if (p->sched_cap[slot] < B) {
        cudaFree(p->d_sched_index[slot]);
        p->d_sched_index[slot] = nullptr;
        p->sched_cap[slot] = 0;
        // idx_a | idx_b | key_a | key_b | counts [256][PV_SCHED_MAX_TILES] | digit_total [256] | digit_base [256] | sums [2]
        CUDA_OK(cudaMalloc(&p->d_sched_index[slot], 4 * per * sizeof(int32_t) + sizeof(uint32_t) * (256 * PV_SCHED_MAX_TILES + 512) +
                                                       2 * sizeof(long long)));
        p->sched_cap[slot] = B;
    }
    const size_t cap = ((size_t)p->sched_cap[slot] + 63) / 64 * 64;
    pv_sched_params sp;
    sp.P = P;
    sp.B = B;
    sp.n_key_cols = std::min<int>((int)p->col_target.size(), 2);
    sp.idx_a = p->d_sched_index[slot];
    sp.idx_b = sp.idx_a + cap;
    sp.key_a = (uint32_t*)(sp.idx_b + cap);
    sp.key_b = sp.key_a + cap;
    sp.counts = sp.key_b + cap;
    sp.digit_total = sp.counts + 256 * PV_SCHED_MAX_TILES;
    sp.digit_base = sp.digit_total + 256;
    sp.sums = (long long*)(sp.digit_base + 256);
    sp.n_tiles = (int)std::min<int64_t>(PV_SCHED_MAX_TILES, std::max<int64_t>(1, (B + 127) / 128));
    sp.tile = (B + sp.n_tiles - 1) / sp.n_tiles;
    sp.n_tiles = (int)((B + sp.tile - 1) / sp.tile);
    CUDA_OK(cudaMemsetAsync(sp.sums, 0, 2 * sizeof(long long), s));
    const unsigned blocks = (unsigned)std::min<int64_t>((B + 255) / 256, (int64_t)p->num_sms * 8);
    const unsigned tile_blocks = (unsigned)((sp.n_tiles + 3) / 4);
    pv_sched_mean_kernel<<<blocks, 256, 0, s>>>(sp);
    pv_sched_key_kernel<<<blocks, 256, 0, s>>>(sp);
    // pass 1: low digit a -> b; pass 2: high digit b -> a
    pv_sched_hist_kernel<<<tile_blocks, 128, 0, s>>>(sp, sp.key_a, 0);
    pv_sched_rowscan_kernel<<<256, 256, 0, s>>>(sp.counts, sp.n_tiles, sp.digit_total);
    pv_sched_digitscan_kernel<<<1, 32, 0, s>>>(sp.digit_total, sp.digit_base);
    pv_sched_scatter_kernel<<<tile_blocks, 128, 0, s>>>(sp, sp.key_a, sp.idx_a, sp.key_b, sp.idx_b, 0);
    pv_sched_hist_kernel<<<tile_blocks, 128, 0, s>>>(sp, sp.key_b, 8);
    pv_sched_rowscan_kernel<<<256, 256, 0, s>>>(sp.counts, sp.n_tiles, sp.digit_total);
    pv_sched_digitscan_kernel<<<1, 32, 0, s>>>(sp.digit_total, sp.digit_base);
    pv_sched_scatter_kernel<<<tile_blocks, 128, 0, s>>>(sp, sp.key_b, sp.idx_b, sp.key_a, sp.idx_a, 8);
    g_launches.fetch_add(10);
    CUDA_OK(cudaGetLastError());
    *index_out = sp.idx_a;
    return 0;
}

static int launch(pv_problem* p, int mode, int64_t B, const double* P, double* Y, double* S, double* logp,
                  double* grad, int32_t* status, int32_t* steps, cudaStream_t s, double* fim = nullptr, int slot = 0) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (B < 0) return fail(PV_E_ARG, "B < 0");
    if (B == 0) return 0;
    if (B >= ((int64_t)1 << 31)) return fail(PV_E_ARG, "B too large for one launch (< 2^31)");
    if (p->tgrid.empty()) return fail(PV_E_STATE, "timepoints not set");
    const bool need_logp = (mode & PV_MODE_LOGP) != 0;
    const bool sens = (mode & PV_MODE_SENS) != 0;
    if (p->obs_idx.empty()) return fail(PV_E_STATE, "observables not set");
    if (need_logp && p->stats.empty()) return fail(PV_E_STATE, "data not set");
    if ((int64_t)p->tgrid.size() * (int64_t)p->obs_idx.size() * (int64_t)std::max(p->n_data, 1) >= ((int64_t)1 << 29))
        return fail(PV_E_ARG, "T * n_obs * n_data too large (< 2^29)");
    if (!p->col_target.empty() && !P) return fail(PV_E_ARG, "P is NULL but columns are configured");
    if (sens && p->tab->n_sens == 0) return fail(PV_E_MODEL, "model was generated without sensitivity parameters");
    if ((mode & PV_MODE_TRAJ) && !Y) return fail(PV_E_ARG, "Y is NULL");
    if (mode == PV_MODE_TRAJ_SENS && !S) return fail(PV_E_ARG, "S is NULL");
    if (need_logp && !logp) return fail(PV_E_ARG, "logp is NULL");
    if ((mode & PV_MODE_LOGP) && (mode & PV_MODE_SENS) && !grad) return fail(PV_E_ARG, "grad is NULL");
    if ((mode & PV_MODE_FIM) && !fim) return fail(PV_E_ARG, "fim is NULL");
    CUDA_OK(cudaSetDevice(p->device));

    pv_launch_params lp;
    memset(&lp, 0, sizeof(lp));
    lp.B = B;
    lp.P = P;
    lp.tgrid = p->d_tgrid;
    lp.stats = p->d_stats;
    lp.Y = Y;
    lp.S = S;
    lp.logp = logp;
    lp.grad = grad;
    lp.fim = fim;
    lp.status = status;
    lp.steps = steps;
    lp.rtol = p->rtol;
    lp.atol = p->atol;
    lp.n_cols = (int)p->col_target.size();
    for (int j = 0; j < lp.n_cols; ++j) lp.col_target[j] = p->col_target[j];
    for (int j = lp.n_cols; j < PV_MAX_COLS; ++j) lp.col_target[j] = -1;
    lp.T = (int)p->tgrid.size();
    lp.n_obs = (int)p->obs_idx.size();
    for (int k = 0; k < lp.n_obs; ++k) lp.obs_idx[k] = p->obs_idx[k];
    lp.n_data = need_logp ? p->n_data : 1;
    lp.noise_model = p->noise_model;
    lp.max_steps = p->max_steps;
    lp.t0 = p->t0;
    lp.counter = next_counter(p);
    CUDA_OK(cudaMemsetAsync(lp.counter, 0, sizeof(unsigned long long), s));
    int method = p->method;
    if (method == PV_METHOD_AUTO) method = p->tab->stiff ? PV_METHOD_RODAS4 : PV_METHOD_DOPRI5;
    // measured (tools/sweep_schedule*.py): the cheap explicit steps profit from refilling idle lanes early (6.83 ms
    // without refill, 6.20 ms at 6 on the C2 cloud); in the large Rosenbrock kernels the divergent start-up code costs
    // more than the idle lanes (icg logp+grad 119 ms at 6, 95 ms at 32)
    lp.refill_min = p->refill_min > 0 ? p->refill_min : (method == PV_METHOD_RODAS4 ? 32 : 6);
    // explicit first pass with a step budget, implicit second pass for what did not finish (stiff parameter values)
    const bool two_pass = p->method == PV_METHOD_AUTO && !p->tab->stiff && p->switch_steps < p->max_steps;
    lp.stiff_after = 0x7fffffff;
    if (two_pass) {
        if (int rc = ensure_scratch(p, slot, B)) return rc;
        lp.max_steps = p->switch_steps;
        lp.stiff_after = p->stiff_after;
        if (!lp.status) lp.status = p->d_scratch_status[slot];
    }
    // auto: from 32 768 trajectories - below that a launch is a wave or two, its duration is its longest trajectory, and the
    // sort's ten small launches only add latency (C3, 8000 trajectories: 0.67 ms unsorted, 0.76 ms sorted)
    if (!p->col_target.empty() && (p->locality == 2 || (p->locality == 0 && B >= 32768)))
        if (int rc = locality_schedule(p, slot, B, P, s, &lp.index_map)) return rc;
    size_t shmem = sizeof(double) * (lp.T + p->tab->n_theta + p->tab->n_states);
    if (need_logp && lp.n_data == 1) shmem += sizeof(double) * PV_QSTRIDE * lp.T * lp.n_obs;
    if (shmem > (size_t)PV_MAX_DYN_SMEM) return fail(PV_E_ARG, "time grid / data too large for shared memory");
    bool obsid = lp.n_obs == p->tab->n_states;
    for (int k = 0; k < lp.n_obs && obsid; ++k) obsid = (lp.obs_idx[k] == k);
    // lockstep warps (one common step size per warp) pay when the lanes integrate similar trajectories: with the locality
    // schedule.  The stiff re-run pass (RODAS4 on a handful of rows) is not affected.
    const bool sorted = lp.index_map != nullptr;
    const int lockstep = p->lockstep == 2 || (p->lockstep == 0 && sorted && PV_LOCKSTEP_AUTO) ? 1 : 0;
    const pv_launch_ctx cx{p->num_sms, p->layout, p->max_ctas_per_sm, lockstep, p->roles, p->theta_base.data(), p->y0_base.data()};
    cudaError_t e = cudaErrorInvalidValue;
    switch (p->model) {
        PV_MODEL_DISPATCH(e = launch_model, (method, mode, obsid, lp, cx, shmem, s))
    }
    if (e != cudaSuccess) return fail(PV_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    if (two_pass) {
        int* cnt = p->d_scratch_count[slot];
        CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int), s));
        pv_compact_status_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(lp.status, B, (1u << PV_ERR_MAX_STEPS) | (1u << PV_ERR_STIFF),
                                                                           p->d_scratch_index[slot], cnt);
        g_launches.fetch_add(1);
        CUDA_OK(cudaGetLastError());
        lp.max_steps = p->max_steps;
        lp.refill_min = p->refill_min > 0 ? p->refill_min : 32;
        lp.index_map = p->d_scratch_index[slot];
        lp.n_items_dev = cnt;
        lp.counter = next_counter(p);
        CUDA_OK(cudaMemsetAsync(lp.counter, 0, sizeof(unsigned long long), s));
        e = cudaErrorInvalidValue;
        switch (p->model) {
            PV_MODEL_DISPATCH(e = launch_model, (PV_METHOD_RODAS4, mode, false, lp, cx, shmem, s))
        }
        if (e != cudaSuccess) return fail(PV_E_CUDA, std::string("kernel launch (stiff re-run): ") + cudaGetErrorString(e));
    }
    return 0;
}

static int segment_sum(const double* x, int64_t B, int64_t seg_len, double* out, cudaStream_t s) {
    if (seg_len <= 0 || !out) return 0;
    if (B % seg_len) return fail(PV_E_ARG, "B must be a multiple of seg_len");
    const int64_t n_seg = B / seg_len;
    const int warps = 4;
    pv_segment_sum_kernel<<<(unsigned)((n_seg + warps - 1) / warps), warps * 32, 0, s>>>(x, seg_len, n_seg, out);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int pv_simulate(pv_problem* p, int64_t B, const double* P_dev, double* Y_dev, int32_t* status_dev,
                           int32_t* steps_dev, void* stream) {
    return launch(p, PV_MODE_TRAJ, B, P_dev, Y_dev, nullptr, nullptr, nullptr, status_dev, steps_dev, (cudaStream_t)stream);
}

extern "C" int pv_simulate_logp(pv_problem* p, int64_t B, const double* P_dev, double* Y_dev, double* logp_dev,
                                int32_t* status_dev, int32_t* steps_dev, void* stream) {
    return launch(p, PV_MODE_TRAJ_LOGP, B, P_dev, Y_dev, nullptr, logp_dev, nullptr, status_dev, steps_dev, (cudaStream_t)stream);
}

extern "C" int pv_simulate_sens(pv_problem* p, int64_t B, const double* P_dev, double* Y_dev, double* S_dev,
                                int32_t* status_dev, int32_t* steps_dev, void* stream) {
    return launch(p, PV_MODE_TRAJ_SENS, B, P_dev, Y_dev, S_dev, nullptr, nullptr, status_dev, steps_dev, (cudaStream_t)stream);
}

extern "C" int pv_logp(pv_problem* p, int64_t B, const double* P_dev, double* logp_dev, int64_t seg_len,
                       double* logp_seg_dev, int32_t* status_dev, int32_t* steps_dev, void* stream) {
    if (int rc = launch(p, PV_MODE_LOGP, B, P_dev, nullptr, nullptr, logp_dev, nullptr, status_dev, steps_dev, (cudaStream_t)stream)) return rc;
    return segment_sum(logp_dev, B, seg_len, logp_seg_dev, (cudaStream_t)stream);
}

extern "C" int pv_logp_grad(pv_problem* p, int64_t B, const double* P_dev, double* logp_dev, double* grad_dev,
                            int64_t seg_len, double* logp_seg_dev, int32_t* status_dev, int32_t* steps_dev, void* stream) {
    if (int rc = launch(p, PV_MODE_LOGP_GRAD, B, P_dev, nullptr, nullptr, logp_dev, grad_dev, status_dev, steps_dev, (cudaStream_t)stream)) return rc;
    return segment_sum(logp_dev, B, seg_len, logp_seg_dev, (cudaStream_t)stream);
}

extern "C" int pv_logp_grad_fim(pv_problem* p, int64_t B, const double* P_dev, double* logp_dev, double* grad_dev,
                                double* fim_dev, int32_t* status_dev, int32_t* steps_dev, void* stream) {
    return launch(p, PV_MODE_LOGP_GRAD_FIM, B, P_dev, nullptr, nullptr, logp_dev, grad_dev, status_dev, steps_dev,
                  (cudaStream_t)stream, fim_dev);
}

// ---------------------------------------------------------------------------------------------------------
// hierarchical objective, device resident (pv_hier.cuh)
// ---------------------------------------------------------------------------------------------------------
extern "C" int pv_hier_logp_grad(pv_problem* p, int n_chains, int64_t n_local, const double* theta_dev, const double* mu_dev,
                                 const double* omega_dev, double* dtheta_dev, double* red_dev, int32_t* status_dev,
                                 int32_t* steps_dev, void* stream) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (n_chains <= 0 || n_local < 0) return fail(PV_E_ARG, "n_chains > 0 and n_local >= 0 required");
    if (!mu_dev || !omega_dev || !red_dev) return fail(PV_E_ARG, "mu / omega / red is NULL");
    const int P = (int)p->col_target.size();
    if (P < 1 || P > PV_HIER_MAX_P) return fail(PV_E_STATE, "configure 1..8 per-individual columns first");
    const int64_t B = (int64_t)n_chains * n_local;
    if (B > 0 && (!theta_dev || !dtheta_dev)) return fail(PV_E_ARG, "theta / dtheta is NULL");
    CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    pv_hier_params hp;
    memset(&hp, 0, sizeof(hp));
    for (int j = 0; j < P; ++j) {
        int row = -1;
        for (int k = 0; k < p->tab->n_sens; ++k)
            if (p->tab->sens_index[k] == p->col_target[j]) row = k;
        if (row < 0)
            return fail(PV_E_MODEL, std::string("column '") + p->tab->theta_ids[std::min(p->col_target[j], p->tab->n_theta - 1)] +
                                        "' is not a compiled sensitivity parameter of model '" + p->tab->name + "'");
        hp.sens_row[j] = row;
    }
    const int ns = p->tab->n_sens;
    const size_t need = sizeof(double) * (size_t)(1 + ns) * (size_t)std::max<int64_t>(B, 1);
    if (p->hier_bytes < need) {       // grows on the first call of a given size only
        cudaFree(p->d_hier);
        p->d_hier = nullptr;
        p->hier_bytes = 0;
        CUDA_OK(cudaMalloc(&p->d_hier, need));
        p->hier_bytes = need;
    }
    double* d_logp = p->d_hier;
    double* d_grad = p->d_hier + std::max<int64_t>(B, 1);
    if (B > 0)
        if (int rc = launch(p, PV_MODE_LOGP_GRAD, B, theta_dev, nullptr, nullptr, d_logp, d_grad, status_dev, steps_dev, s)) return rc;
    hp.n_chains = n_chains;
    hp.n_local = n_local;
    hp.n_par = P;
    hp.theta = theta_dev;
    hp.mu = mu_dev;
    hp.omega = omega_dev;
    hp.logp = d_logp;
    hp.grad = d_grad;
    hp.dtheta = dtheta_dev;
    hp.red = red_dev;
    pv_hier_epilogue_kernel<<<n_chains, PV_HIER_THREADS, 0, s>>>(hp);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int pv_hier_finish(int n_chains, int n_par, double* red_dev, const double* mu_dev, const double* omega_dev,
                              const double* mu_prior_mean_host, const double* mu_prior_sd_host,
                              const double* omega_prior_host, void* stream) {
    if (n_chains <= 0 || n_par < 1 || n_par > PV_HIER_MAX_P) return fail(PV_E_ARG, "n_chains > 0 and 1 <= n_par <= 8 required");
    if (!red_dev || !mu_dev || !omega_dev || !mu_prior_mean_host || !mu_prior_sd_host || !omega_prior_host)
        return fail(PV_E_ARG, "NULL argument");
    pv_hier_prior pr;
    memset(&pr, 0, sizeof(pr));
    for (int q = 0; q < n_par; ++q) {
        if (!(mu_prior_sd_host[q] > 0) || !(omega_prior_host[q] > 0)) return fail(PV_E_ARG, "prior scales must be > 0");
        pr.m[q] = mu_prior_mean_host[q];
        pr.s[q] = mu_prior_sd_host[q];
        pr.h[q] = omega_prior_host[q];
    }
    pv_hier_finish_kernel<<<(n_chains + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n_chains, n_par, red_dev, mu_dev, omega_dev, pr);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Hamiltonian Monte Carlo on the hierarchical posterior, device resident (pv_hmc.cuh)
// ---------------------------------------------------------------------------------------------------------
static int hmc_check(const pv_hmc_state* s) {
    if (!s) return fail(PV_E_ARG, "pv_hmc: state is NULL");
    if (s->n_chains <= 0 || s->n_par < 1 || s->n_par > PV_HIER_MAX_P || s->n_local < 0 || s->n_total < s->n_local || s->first < 0 ||
        s->first + s->n_local > s->n_total)
        return fail(PV_E_ARG, "pv_hmc: n_chains > 0, 1 <= n_par <= 8, 0 <= first, first + n_local <= n_total required");
    const void* need[] = {s->lth, s->mu, s->lom, s->g_lth, s->g_hyp, s->lp, s->lth_q, s->mu_q, s->lom_q, s->g_lth_q, s->g_hyp_q,
                          s->p_lth, s->p_hyp, s->im_lth, s->im_hyp, s->eps, s->theta, s->omega, s->dtheta, s->red, s->energy,
                          s->ok, s->da, s->w_mean_lth, s->w_m2_lth, s->w_mean_hyp, s->w_m2_hyp, s->samples_lp, s->acc_sum, s->n_div};
    for (const void* q : need)
        if (!q) return fail(PV_E_ARG, "pv_hmc: NULL buffer in the state (only samples_lth / samples_hyp may be NULL)");
    return 0;
}
extern "C" int pv_hmc_begin(const pv_hmc_state* s, const double* z_dev, void* stream) {
    if (int rc = hmc_check(s)) return rc;
    if (!z_dev) return fail(PV_E_ARG, "pv_hmc_begin: z is NULL");
    pv_hmc_begin_kernel<<<s->n_chains, PV_HMC_THREADS, 0, (cudaStream_t)stream>>>(*s, z_dev);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int pv_hmc_drift(const pv_hmc_state* s, double scale, void* stream) {
    if (int rc = hmc_check(s)) return rc;
    pv_hmc_drift_kernel<<<s->n_chains, PV_HMC_THREADS, 0, (cudaStream_t)stream>>>(*s, scale);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int pv_hmc_kick(const pv_hmc_state* s, double scale, int last, void* stream) {
    if (int rc = hmc_check(s)) return rc;
    pv_hmc_kick_kernel<<<s->n_chains, PV_HMC_THREADS, 0, (cudaStream_t)stream>>>(*s, scale, last);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int pv_hmc_accept(const pv_hmc_state* s, int mode, int64_t it, int64_t n_warmup, const double* u_dev, double target_accept,
                             double max_energy_error, int adapt_mass, int64_t w_n, void* stream) {
    if (int rc = hmc_check(s)) return rc;
    if (mode < 0 || mode > 2 || (mode != 0 && !u_dev) || (mode == 2 && it < n_warmup) || it < 0)
        return fail(PV_E_ARG, "pv_hmc_accept: mode 0..2, u for modes 1 / 2, it >= n_warmup for mode 2 required");
    pv_hmc_accept_kernel<<<s->n_chains, PV_HMC_THREADS, 0, (cudaStream_t)stream>>>(*s, mode, it, n_warmup, u_dev, target_accept,
                                                                                   max_energy_error, adapt_mass, w_n);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// No-U-Turn sampler on the hierarchical posterior, device resident (pv_nuts.cuh)
// ---------------------------------------------------------------------------------------------------------
static int nuts_check(const pv_nuts_state* s) {
    if (!s) return fail(PV_E_ARG, "pv_nuts: state is NULL");
    if (s->n_chains <= 0 || s->n_chains > PV_NUTS_MAX_CHAINS || s->n_par < 1 || s->n_par > PV_HIER_MAX_P || s->n_ind < 0 ||
        s->max_depth < 1 || s->max_depth > 20 || s->n_samples < 0)
        return fail(PV_E_ARG, "pv_nuts: 1 <= n_chains <= 64, 1 <= n_par <= 8, 1 <= max_depth <= 20 required");
    const void* need[] = {s->vec, s->ck, s->sc, s->ic, s->theta, s->mu_q, s->omega, s->dtheta, s->red, s->samples, s->samples_lp,
                          s->samples_depth};
    for (const void* q : need)
        if (!q) return fail(PV_E_ARG, "pv_nuts: NULL buffer in the state");
    return 0;
}
#define PV_NUTS_LAUNCH(kernel, ...)                                                                  \
    kernel<<<s->n_chains, PV_NUTS_THREADS, 0, (cudaStream_t)stream>>>(*s, ##__VA_ARGS__);            \
    g_launches.fetch_add(1);                                                                         \
    CUDA_OK(cudaGetLastError());                                                                     \
    return 0;
extern "C" int pv_nuts_begin(const pv_nuts_state* s, const double* z_dev, void* stream) {
    if (int rc = nuts_check(s)) return rc;
    if (!z_dev) return fail(PV_E_ARG, "pv_nuts_begin: z is NULL");
    PV_NUTS_LAUNCH(pv_nuts_begin_kernel, z_dev)
}
extern "C" int pv_nuts_start(const pv_nuts_state* s, const pv_nuts_uniforms* u, void* stream) {
    if (int rc = nuts_check(s)) return rc;
    if (!u) return fail(PV_E_ARG, "pv_nuts_start: uniforms are NULL");
    PV_NUTS_LAUNCH(pv_nuts_start_kernel, *u)
}
extern "C" int pv_nuts_drift(const pv_nuts_state* s, void* stream) {
    if (int rc = nuts_check(s)) return rc;
    PV_NUTS_LAUNCH(pv_nuts_drift_kernel)
}
extern "C" int pv_nuts_leaf(const pv_nuts_state* s, const pv_nuts_uniforms* u, double max_energy_error, void* stream) {
    if (int rc = nuts_check(s)) return rc;
    if (!u) return fail(PV_E_ARG, "pv_nuts_leaf: uniforms are NULL");
    PV_NUTS_LAUNCH(pv_nuts_leaf_kernel, *u, max_energy_error)
}
extern "C" int pv_nuts_merge(const pv_nuts_state* s, const pv_nuts_uniforms* u, void* stream) {
    if (int rc = nuts_check(s)) return rc;
    if (!u) return fail(PV_E_ARG, "pv_nuts_merge: uniforms are NULL");
    PV_NUTS_LAUNCH(pv_nuts_merge_kernel, *u)
}
extern "C" int pv_nuts_finish(const pv_nuts_state* s, int mode, int64_t it, int64_t n_warmup, double target_accept, int adapt_mass,
                              int64_t w_n, void* stream) {
    if (int rc = nuts_check(s)) return rc;
    if (mode < 0 || mode > 2 || it < 0 || (mode == 2 && (it < n_warmup || it - n_warmup >= s->n_samples)))
        return fail(PV_E_ARG, "pv_nuts_finish: mode 0..2, n_warmup <= it < n_warmup + n_samples for mode 2 required");
    PV_NUTS_LAUNCH(pv_nuts_finish_kernel, mode, it, n_warmup, target_accept, adapt_mass, w_n)
}
#undef PV_NUTS_LAUNCH

// ---------------------------------------------------------------------------------------------------------
// host-buffer pipelines
// ---------------------------------------------------------------------------------------------------------
static int ensure_dev(pv_problem* p, int slot, size_t bytes) {
    if (p->stage_bytes[slot] >= bytes) return 0;
    cudaFree(p->stage_dev[slot]);
    p->stage_dev[slot] = nullptr;
    p->stage_bytes[slot] = 0;
    CUDA_OK(cudaMalloc(&p->stage_dev[slot], bytes));
    p->stage_bytes[slot] = bytes;
    return 0;
}

extern "C" int pv_simulate_host(pv_problem* p, int64_t B, const double* P_host, double* Y_host, double* logp_host,
                                int32_t* status_host, int32_t* steps_host) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (B <= 0) return B == 0 ? 0 : fail(PV_E_ARG, "B < 0");
    if (p->tgrid.empty() || p->obs_idx.empty()) return fail(PV_E_STATE, "timepoints / observables not set");
    if (!Y_host) return fail(PV_E_ARG, "Y_host is NULL");
    if (logp_host && p->stats.empty()) return fail(PV_E_STATE, "logp requested but data not set");
    CUDA_OK(cudaSetDevice(p->device));
    const int n_cols = (int)p->col_target.size();
    const int64_t rows = (int64_t)p->tgrid.size() * (int64_t)p->obs_idx.size();
    // chunks sized so that copy-out of chunk i overlaps the kernel of chunk i+1
    // measured on the C2 workload (tools/time_e2e.py; raw pinned D2H of this box: 57 GB/s, one unchunked copy of the
    // results alone: 21.9 ms): 16 Ki rows 23.6 ms, 32 Ki 23.0 ms, 64 Ki 23.4 ms, 128 Ki 24.4 ms per 10^6 trajectories
    const int64_t chunk = std::min<int64_t>(B, p->host_chunk_rows);   // pv_problem_set_tuning
    // per slot: P [n_cols][chunk], Y [chunk][rows], logp [chunk], status [chunk] + steps [2][chunk] (int32)
    const size_t bytesP = sizeof(double) * (size_t)std::max(n_cols, 1) * chunk;
    const size_t bytesY = sizeof(double) * (size_t)rows * chunk;
    const size_t bytesL = sizeof(double) * (size_t)chunk;
    const size_t bytesI = sizeof(int32_t) * 3 * (size_t)chunk;
    const size_t slot_bytes = bytesP + bytesY + bytesL + bytesI;
    const int n_slots = (B > chunk) ? 3 : 1;
    for (int s = 0; s < n_slots; ++s)
        if (int rc = ensure_dev(p, s, slot_bytes)) return rc;
    CUDA_OK(cudaMemsetAsync(p->d_failed, 0, sizeof(int), p->streams[0]));
    CUDA_OK(cudaStreamSynchronize(p->streams[0]));
    CUDA_OK(cudaEventRecord(p->ev0, p->streams[0]));
    // Slot pipeline: H2D + kernels of chunk k run on the slot's own stream; every device->host copy goes through ONE copy
    // stream in chunk order (events hand the slot over and back), so the copy engine drains the results back to back
    // instead of arbitrating between three streams.
    cudaStream_t cs = p->copy_stream;
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++k) {
        const int slot = k % n_slots;
        cudaStream_t s = p->streams[slot];
        const int64_t nb = std::min(chunk, B - b0);
        double* dP = p->stage_dev[slot];
        double* dY = (double*)((char*)dP + bytesP);
        double* dL = (double*)((char*)dY + bytesY);
        int32_t* dStatus = (int32_t*)((char*)dL + bytesL);
        int32_t* dSteps = dStatus + chunk;
        if (k >= n_slots) CUDA_OK(cudaStreamWaitEvent(s, p->ev_drained[slot], 0));   // slot buffers are free again
        if (n_cols > 0)
            CUDA_OK(cudaMemcpy2DAsync(dP, sizeof(double) * nb, P_host + b0, sizeof(double) * B, sizeof(double) * nb, n_cols,
                                      cudaMemcpyHostToDevice, s));
        if (int rc = launch(p, logp_host ? PV_MODE_TRAJ_LOGP : PV_MODE_TRAJ, nb, n_cols ? dP : nullptr, dY, nullptr,
                            logp_host ? dL : nullptr, nullptr, dStatus, dSteps, s, nullptr, 1 + slot))
            return rc;
        pv_count_failed_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(dStatus, nb, p->d_failed);
        g_launches.fetch_add(1);
        CUDA_OK(cudaEventRecord(p->ev_ready[slot], s));
        CUDA_OK(cudaStreamWaitEvent(cs, p->ev_ready[slot], 0));
        CUDA_OK(cudaMemcpyAsync(Y_host + b0 * rows, dY, sizeof(double) * rows * nb, cudaMemcpyDeviceToHost, cs));
        if (logp_host) CUDA_OK(cudaMemcpyAsync(logp_host + b0, dL, sizeof(double) * nb, cudaMemcpyDeviceToHost, cs));
        if (status_host) CUDA_OK(cudaMemcpyAsync(status_host + b0, dStatus, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, cs));
        if (steps_host)
            CUDA_OK(cudaMemcpy2DAsync(steps_host + b0, sizeof(int32_t) * B, dSteps, sizeof(int32_t) * nb, sizeof(int32_t) * nb, 2,
                                      cudaMemcpyDeviceToHost, cs));
        CUDA_OK(cudaEventRecord(p->ev_drained[slot], cs));
    }
    CUDA_OK(cudaStreamSynchronize(cs));
    for (int s = 0; s < n_slots; ++s) CUDA_OK(cudaStreamSynchronize(p->streams[s]));
    int failed = 0;
    CUDA_OK(cudaMemcpy(&failed, p->d_failed, sizeof(int), cudaMemcpyDeviceToHost));
    return failed;
}

static int logp_host_impl(pv_problem* p, bool with_grad, int64_t B, const double* P_host, double* logp_host,
                          double* grad_host, int64_t seg_len, double* logp_seg_host, double* fim_host = nullptr) {
    if (!p) return fail(PV_E_ARG, "problem is NULL");
    if (B <= 0) return B == 0 ? 0 : fail(PV_E_ARG, "B < 0");
    if (seg_len > 0 && (B % seg_len)) return fail(PV_E_ARG, "B must be a multiple of seg_len");
    CUDA_OK(cudaSetDevice(p->device));
    const int n_cols = (int)p->col_target.size();
    const int ns = p->tab->n_sens;
    const int64_t n_seg = seg_len > 0 ? B / seg_len : 0;
    const size_t bytesP = sizeof(double) * (size_t)std::max(n_cols, 1) * B;
    const size_t bytesL = sizeof(double) * (size_t)B;
    const size_t bytesG = with_grad ? sizeof(double) * (size_t)ns * B : 0;
    const size_t bytesF = fim_host ? sizeof(double) * (size_t)(ns * (ns + 1) / 2) * B : 0;
    const size_t bytesS = sizeof(double) * (size_t)std::max<int64_t>(n_seg, 1);
    const size_t bytesI = sizeof(int32_t) * (size_t)B;
    if (int rc = ensure_dev(p, 0, bytesP + bytesL + bytesG + bytesF + bytesS + bytesI)) return rc;
    cudaStream_t s = p->streams[0];
    double* dP = p->stage_dev[0];
    double* dL = (double*)((char*)dP + bytesP);
    double* dG = (double*)((char*)dL + bytesL);
    double* dF = (double*)((char*)dG + bytesG);
    double* dS = (double*)((char*)dF + bytesF);
    int32_t* dStatus = (int32_t*)((char*)dS + bytesS);
    CUDA_OK(cudaMemsetAsync(p->d_failed, 0, sizeof(int), s));
    if (n_cols > 0) CUDA_OK(cudaMemcpyAsync(dP, P_host, sizeof(double) * (size_t)n_cols * B, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaEventRecord(p->ev0, s));
    int rc;
    if (fim_host) {
        rc = pv_logp_grad_fim(p, B, n_cols ? dP : nullptr, dL, dG, dF, dStatus, nullptr, s);
        if (!rc) rc = segment_sum(dL, B, seg_len, n_seg ? dS : nullptr, s);
    } else {
        rc = with_grad ? pv_logp_grad(p, B, n_cols ? dP : nullptr, dL, dG, seg_len, n_seg ? dS : nullptr, dStatus, nullptr, s)
                       : pv_logp(p, B, n_cols ? dP : nullptr, dL, seg_len, n_seg ? dS : nullptr, dStatus, nullptr, s);
    }
    if (rc) return rc;
    CUDA_OK(cudaEventRecord(p->ev1, s));
    pv_count_failed_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(dStatus, B, p->d_failed);
    g_launches.fetch_add(1);
    if (logp_host) CUDA_OK(cudaMemcpyAsync(logp_host, dL, bytesL, cudaMemcpyDeviceToHost, s));
    if (with_grad && grad_host) CUDA_OK(cudaMemcpyAsync(grad_host, dG, bytesG, cudaMemcpyDeviceToHost, s));
    if (fim_host) CUDA_OK(cudaMemcpyAsync(fim_host, dF, bytesF, cudaMemcpyDeviceToHost, s));
    if (n_seg && logp_seg_host) CUDA_OK(cudaMemcpyAsync(logp_seg_host, dS, sizeof(double) * n_seg, cudaMemcpyDeviceToHost, s));
    int failed = 0;
    CUDA_OK(cudaMemcpyAsync(&failed, p->d_failed, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->ev0, p->ev1) == cudaSuccess) p->last_kernel_ms = ms;
    return failed;
}

extern "C" int pv_logp_host(pv_problem* p, int64_t B, const double* P_host, double* logp_host, int64_t seg_len,
                            double* logp_seg_host) {
    return logp_host_impl(p, false, B, P_host, logp_host, nullptr, seg_len, logp_seg_host);
}
extern "C" int pv_logp_grad_host(pv_problem* p, int64_t B, const double* P_host, double* logp_host, double* grad_host,
                                 int64_t seg_len, double* logp_seg_host) {
    return logp_host_impl(p, true, B, P_host, logp_host, grad_host, seg_len, logp_seg_host);
}

extern "C" void* pv_host_alloc(size_t bytes) {
    void* ptr = nullptr;
    if (cudaHostAlloc(&ptr, bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(PV_E_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return ptr;
}
extern "C" void pv_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

extern "C" int pv_logp_grad_fim_host(pv_problem* p, int64_t B, const double* P_host, double* logp_host, double* grad_host,
                                     double* fim_host, int64_t seg_len, double* logp_seg_host) {
    if (!fim_host) return fail(PV_E_ARG, "fim_host is NULL");
    return logp_host_impl(p, true, B, P_host, logp_host, grad_host, seg_len, logp_seg_host, fim_host);
}

// ---- host bookkeeping of the batched sampler (pv_sampler_host.cuh) ----------------------------------------------------
extern "C" int64_t pv_sampler_propose(int64_t Q, int C, int dim, const double* cov, const double* x, const double* xi,
                                      const double* lb, const double* ub, double* x_new, uint8_t* inside, double* x_eval) {
    if (Q < 0 || C < 1 || dim < 1 || !cov || !x || !xi || !lb || !ub || !x_new || !inside || !x_eval)
        return fail(PV_E_ARG, "pv_sampler_propose: bad arguments");
    const int64_t r = pv_sampler_propose_impl(Q, C, dim, cov, x, xi, lb, ub, x_new, inside, x_eval);
    if (r < 0) return fail(PV_E_ARG, "pv_sampler_propose: dim too large");
    return r;
}
extern "C" int pv_sampler_step(int64_t Q, int C, int dim, int64_t it, int64_t n_trace, const pv_sampler_options* o, double* x,
                               double* llh, double* lpr, const double* x_new, const double* llh_new, const double* lpr_new,
                               const uint8_t* inside, const double* u_acc, const double* u_swap, double* betas,
                               double* mean_hist, double* cov_hist, double* cov_scale, double* cov, double* n_acc,
                               double* n_swap, double* n_swap_try, double* trace_x, double* trace_lp, double* trace_pr) {
    if (Q < 0 || C < 1 || dim < 1 || it < 1 || it >= n_trace || !o || !x || !llh || !lpr || !x_new || !llh_new || !lpr_new ||
        !inside || !u_acc || (C > 1 && !u_swap) || !betas || !mean_hist || !cov_hist || !cov_scale || !cov || !n_acc ||
        !n_swap || !n_swap_try || !trace_x || !trace_lp || !trace_pr)
        return fail(PV_E_ARG, "pv_sampler_step: bad arguments");
    if (pv_sampler_step_impl(Q, C, dim, it, n_trace, o, x, llh, lpr, x_new, llh_new, lpr_new, inside, u_acc, u_swap, betas,
                             mean_hist, cov_hist, cov_scale, cov, n_acc, n_swap, n_swap_try, trace_x, trace_lp, trace_pr))
        return fail(PV_E_ARG, "pv_sampler_step: dim or number of temperatures too large");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// device-resident sampler (pv_sampler_device.cuh)
// ---------------------------------------------------------------------------------------------------------
struct pv_sampler {
    pv_problem* p = nullptr;
    pv_sampler_dev d;
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t ev_block[2] = {nullptr, nullptr};   // the iterations of the block staged in that half are done
    double *h_xi = nullptr, *h_ua = nullptr, *h_us = nullptr;   // pinned staging, two blocks each
    int64_t blocks_queued = 0;
    double *d_xi = nullptr, *d_ua = nullptr, *d_us = nullptr;
    int32_t* d_status = nullptr;
    double* d_logp = nullptr;
    long long it_host = 0;               // iterations queued so far (the next one to run is it_host + 1)
    int step_threads = 0, step_blocks = 0;
    size_t step_shmem = 0;
};

template <class T>
static T* sampler_alloc(pv_sampler* s, size_t n, bool zero = true) {
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, sizeof(T) * std::max<size_t>(n, 1)) != cudaSuccess) ptr = nullptr;
    if (ptr && zero) cudaMemset(ptr, 0, sizeof(T) * std::max<size_t>(n, 1));
    s->allocs.push_back(ptr);      // a failed allocation stays in the list as NULL: pv_sampler_create checks for it
    return (T*)ptr;
}

// one iteration on the sampler's stream: proposal -> objective (one batched launch for all problems and temperatures) ->
// accept / adapt / swap / trace
static int sampler_enqueue_iteration(pv_sampler* s) {
    const pv_sampler_dev& d = s->d;
    const int64_t QC = d.Q * d.C;
    pv_sampler_propose_kernel<<<(unsigned)((QC + 127) / 128), 128, 0, s->stream>>>(d);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    const int64_t B = QC * d.G;
    if (int rc = launch(s->p, PV_MODE_LOGP, B, d.P, nullptr, nullptr, s->d_logp, nullptr, s->d_status, nullptr, s->stream)) return rc;
    if (int rc = segment_sum(s->d_logp, B, d.G, d.seg, s->stream)) return rc;
    pv_sampler_step_kernel<<<s->step_blocks, s->step_threads, s->step_shmem, s->stream>>>(d);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" void pv_sampler_destroy(pv_sampler* s) {
    if (!s) return;
    if (s->p) cudaSetDevice(s->p->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->exec) cudaGraphExecDestroy(s->exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    for (auto& e : s->ev_block)
        if (e) cudaEventDestroy(e);
    for (void* a : s->allocs) cudaFree(a);
    if (s->h_xi) cudaFreeHost(s->h_xi);
    if (s->h_ua) cudaFreeHost(s->h_ua);
    if (s->h_us) cudaFreeHost(s->h_us);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" pv_sampler* pv_sampler_create(pv_problem* p, int64_t Q, int n_groups, int n_chains, int64_t n_samples,
                                         const pv_sampler_options* options, const double* lb, const double* ub,
                                         const double* prior_loc, const double* prior_scale, const double* x0,
                                         const double* cov0, int64_t rng_block, int use_graph) {
    auto bad = [](const char* m) -> pv_sampler* {
        fail(PV_E_ARG, std::string("pv_sampler_create: ") + m);
        return nullptr;
    };
    if (!p || !options || !lb || !ub || !prior_loc || !prior_scale || !x0) return bad("NULL argument");
    const int NP = (int)p->col_target.size(), G = n_groups, C = n_chains;
    if (Q < 1 || G < 1 || C < 1 || C > 64 || n_samples < 1 || rng_block < 1) return bad("Q, n_groups, n_chains (<= 64), n_samples, rng_block must be >= 1");
    const int dim = NP * G;
    if (NP < 1 || dim > pv_sampler_math::MAXD) return bad("configure the per-group parameters as columns first (n_cols * n_groups <= 16)");
    if (p->stats.empty() || p->n_data != Q * G) return bad("the problem must hold one data set per (problem, group): n_data == Q * n_groups");
    if (cudaSetDevice(p->device) != cudaSuccess) return bad("cudaSetDevice failed");
    auto* s = new pv_sampler();
    s->p = p;
    memset(&s->d, 0, sizeof(s->d));
    pv_sampler_dev& d = s->d;
    d.Q = Q; d.C = C; d.G = G; d.NP = NP; d.dim = dim; d.S = C > 1 ? C - 1 : 1;
    d.n_trace = n_samples + 1;
    d.rng_block = rng_block;
    d.o = *options;
    const int64_t QC = Q * C;
    bool ok = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess;
    s->own_stream = ok;
    for (auto& e : s->ev_block) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    d.x = sampler_alloc<double>(s, QC * dim); d.llh = sampler_alloc<double>(s, QC); d.lpr = sampler_alloc<double>(s, QC);
    d.betas = sampler_alloc<double>(s, QC); d.mean_hist = sampler_alloc<double>(s, QC * dim);
    d.cov_hist = sampler_alloc<double>(s, QC * dim * dim); d.cov_scale = sampler_alloc<double>(s, QC);
    d.cov = sampler_alloc<double>(s, QC * dim * dim); d.n_acc = sampler_alloc<double>(s, QC);
    d.n_swap = sampler_alloc<double>(s, Q * d.S); d.n_swap_try = sampler_alloc<double>(s, Q * d.S);
    d.x_new = sampler_alloc<double>(s, QC * dim); d.lpr_new = sampler_alloc<double>(s, QC);
    d.inside = sampler_alloc<unsigned char>(s, QC);
    d.P = sampler_alloc<double>(s, (size_t)NP * QC * G); d.seg = sampler_alloc<double>(s, QC);
    double* d_lb = sampler_alloc<double>(s, dim); double* d_ub = sampler_alloc<double>(s, dim);
    double* d_pl = sampler_alloc<double>(s, Q * dim); double* d_ps = sampler_alloc<double>(s, Q * dim);
    s->d_xi = sampler_alloc<double>(s, (size_t)rng_block * QC * dim); s->d_ua = sampler_alloc<double>(s, (size_t)rng_block * QC);
    s->d_us = sampler_alloc<double>(s, (size_t)rng_block * Q * d.S);
    d.trace_x = sampler_alloc<double>(s, (size_t)d.n_trace * QC * dim, false);
    d.trace_lp = sampler_alloc<double>(s, (size_t)d.n_trace * QC, false);
    d.trace_pr = sampler_alloc<double>(s, (size_t)d.n_trace * QC, false);
    d.it = sampler_alloc<long long>(s, 1); d.done = sampler_alloc<unsigned int>(s, 1);
    d.n_bad_cov = sampler_alloc<unsigned long long>(s, 1);
    s->d_status = sampler_alloc<int32_t>(s, QC * G); s->d_logp = sampler_alloc<double>(s, QC * G);
    unsigned long long* counters = sampler_alloc<unsigned long long>(s, PV_SAMPLER_COUNTERS);
    for (void* a : s->allocs) ok = ok && a != nullptr;
    ok = ok && cudaHostAlloc(&s->h_xi, sizeof(double) * 2 * rng_block * QC * dim, cudaHostAllocDefault) == cudaSuccess &&
         cudaHostAlloc(&s->h_ua, sizeof(double) * 2 * rng_block * QC, cudaHostAllocDefault) == cudaSuccess &&
         cudaHostAlloc(&s->h_us, sizeof(double) * 2 * rng_block * Q * d.S, cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        fail(PV_E_CUDA, std::string("pv_sampler_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        pv_sampler_destroy(s);
        return nullptr;
    }
    d.lb = d_lb; d.ub = d_ub; d.prior_loc = d_pl; d.prior_scale = d_ps;
    d.xi = s->d_xi; d.u_acc = s->d_ua; d.u_swap = s->d_us;
    // ---- initial state (optimization/sampler.py::sample): every chain of a problem starts at x0[q]; ladder; covariance --
    std::vector<double> hx((size_t)QC * dim), hb(QC), hc((size_t)QC * dim * dim), ones(QC, 1.0);
    for (int64_t q = 0; q < Q; ++q)
        for (int c = 0; c < C; ++c) {
            for (int i = 0; i < dim; ++i) hx[(q * C + c) * dim + i] = x0[q * dim + i];
            // near-exponential temperature ladder: T_c = 1 + (c / (C - 1))^4 (5e4 - 1)
            const double frac = C > 1 ? (double)c / (double)(C - 1) : 0.0;
            hb[q * C + c] = 1.0 / (1.0 + frac * frac * frac * frac * (5e4 - 1.0));
            double* cv = hc.data() + (q * C + c) * dim * dim;
            for (int i = 0; i < dim; ++i)
                for (int j = 0; j < dim; ++j) cv[i * dim + j] = cov0 ? cov0[i * dim + j] : (i == j ? 1.0 : 0.0);
            pv_sampler_math::regularize(dim, cv, options->reg_factor);
        }
    auto up = [&](double* dst, const double* src, size_t n) { return cudaMemcpy(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice) == cudaSuccess; };
    ok = up(d.x, hx.data(), hx.size()) && up(d.mean_hist, hx.data(), hx.size()) && up(d.betas, hb.data(), hb.size()) &&
         up(d.cov, hc.data(), hc.size()) && up(d.cov_hist, hc.data(), hc.size()) && up(d.cov_scale, ones.data(), ones.size()) &&
         up(d_lb, lb, dim) && up(d_ub, ub, dim) && up(d_pl, prior_loc, (size_t)Q * dim) && up(d_ps, prior_scale, (size_t)Q * dim) &&
         cudaStreamSynchronize(0) == cudaSuccess;
    // step kernel geometry: all temperatures of a problem in one block
    const int qpb = std::max(1, 128 / C);
    s->step_threads = qpb * C;
    s->step_blocks = (int)((Q + qpb - 1) / qpb);
    s->step_shmem = (size_t)qpb * d.S;
    // ---- iteration 0 = the objective at the start points (also warms every cache the capture below must not touch) ----
    int rc = ok ? sampler_enqueue_iteration(s) : PV_E_CUDA;
    if (!rc && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = PV_E_CUDA;
    if (!rc && use_graph) {
        p->own_counters = counters;
        p->own_counter_next = 0;
        cudaError_t e = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = sampler_enqueue_iteration(s);
            e = cudaStreamEndCapture(s->stream, &s->graph);
            if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&s->exec, s->graph, 0);
        }
        p->own_counters = nullptr;
        if (e != cudaSuccess && !rc) {
            fail(PV_E_CUDA, std::string("pv_sampler_create: graph capture: ") + cudaGetErrorString(e));
            rc = PV_E_CUDA;
        }
    }
    if (rc) {
        if (ok && rc == PV_E_CUDA && g_last_error.empty()) fail(PV_E_CUDA, "pv_sampler_create: CUDA error");
        pv_sampler_destroy(s);
        return nullptr;
    }
    return s;
}

// queue n_iter iterations; the random numbers of the block are either copied from the caller's arrays or drawn here
static int sampler_advance_impl(pv_sampler* s, int64_t n_iter, const double* xi_host, const double* u_acc_host,
                                const double* u_swap_host, pv_mt19937_state* rng) {
    const pv_sampler_dev& d = s->d;
    if (n_iter < 1 || n_iter > d.rng_block) return fail(PV_E_ARG, "pv_sampler_advance: 1 <= n_iter <= rng_block required");
    if (s->it_host % d.rng_block) return fail(PV_E_STATE, "pv_sampler_advance: only the last block may be shorter than rng_block");
    if (s->it_host + n_iter > d.n_trace - 1) return fail(PV_E_ARG, "pv_sampler_advance: more iterations than n_samples");
    CUDA_OK(cudaSetDevice(s->p->device));
    const int64_t QC = d.Q * d.C;
    // double-buffered staging: block k is drawn / copied into half k % 2 while block k - 1 runs; half k % 2 was last read by
    // the upload of block k - 2, which has completed once block k - 2's event has (recorded after its iterations)
    const int half = (int)((s->it_host / d.rng_block) & 1);
    if (s->blocks_queued >= 2) CUDA_OK(cudaEventSynchronize(s->ev_block[half]));
    double* h_xi = s->h_xi + (size_t)half * d.rng_block * QC * d.dim;
    double* h_ua = s->h_ua + (size_t)half * d.rng_block * QC;
    double* h_us = s->h_us + (size_t)half * d.rng_block * d.Q * d.S;
    if (rng) {
        pv_rng::fill_sampler_block(rng, n_iter, d.Q, d.C, d.dim, h_xi, h_ua, h_us);
    } else {
        memcpy(h_xi, xi_host, sizeof(double) * n_iter * QC * d.dim);
        memcpy(h_ua, u_acc_host, sizeof(double) * n_iter * QC);
        if (d.C > 1) memcpy(h_us, u_swap_host, sizeof(double) * n_iter * d.Q * d.S);
    }
    // the device block is single: its upload is ordered behind the previous block's iterations on the stream
    CUDA_OK(cudaMemcpyAsync(s->d_xi, h_xi, sizeof(double) * n_iter * QC * d.dim, cudaMemcpyHostToDevice, s->stream));
    CUDA_OK(cudaMemcpyAsync(s->d_ua, h_ua, sizeof(double) * n_iter * QC, cudaMemcpyHostToDevice, s->stream));
    if (d.C > 1) CUDA_OK(cudaMemcpyAsync(s->d_us, h_us, sizeof(double) * n_iter * d.Q * d.S, cudaMemcpyHostToDevice, s->stream));
    for (int64_t k = 0; k < n_iter; ++k) {
        if (s->exec) {
            CUDA_OK(cudaGraphLaunch(s->exec, s->stream));
            g_launches.fetch_add(6);     // nodes of the recorded iteration: propose, 2 objective passes + compaction, segment sum, step
        } else if (int rc = sampler_enqueue_iteration(s)) {
            return rc;
        }
    }
    CUDA_OK(cudaEventRecord(s->ev_block[half], s->stream));
    s->it_host += n_iter;
    s->blocks_queued += 1;
    return 0;
}

extern "C" int pv_sampler_advance(pv_sampler* s, int64_t n_iter, const double* xi_host, const double* u_acc_host,
                                  const double* u_swap_host) {
    if (!s || !xi_host || !u_acc_host || (s->d.C > 1 && !u_swap_host)) return fail(PV_E_ARG, "pv_sampler_advance: NULL argument");
    return sampler_advance_impl(s, n_iter, xi_host, u_acc_host, u_swap_host, nullptr);
}

extern "C" int pv_sampler_advance_mt(pv_sampler* s, int64_t n_iter, pv_mt19937_state* rng) {
    if (!s || !rng) return fail(PV_E_ARG, "pv_sampler_advance_mt: NULL argument");
    if (rng->pos < 0 || rng->pos > 624) return fail(PV_E_ARG, "pv_sampler_advance_mt: bad generator state");
    return sampler_advance_impl(s, n_iter, nullptr, nullptr, nullptr, rng);
}

extern "C" int pv_rng_sampler_block(pv_mt19937_state* rng, int64_t n_iter, int64_t Q, int n_chains, int dim, double* xi_host,
                                    double* u_acc_host, double* u_swap_host) {
    if (!rng || n_iter < 0 || Q < 1 || n_chains < 1 || dim < 1 || !xi_host || !u_acc_host || !u_swap_host)
        return fail(PV_E_ARG, "pv_rng_sampler_block: bad arguments");
    if (rng->pos < 0 || rng->pos > 624) return fail(PV_E_ARG, "pv_rng_sampler_block: bad generator state");
    pv_rng::fill_sampler_block(rng, n_iter, Q, n_chains, dim, xi_host, u_acc_host, u_swap_host);
    return 0;
}

extern "C" int pv_sampler_result(pv_sampler* s, double* trace_x_host, double* trace_neglogpost_host, double* trace_neglogprior_host,
                                 double* betas_host, double* n_acc_host, double* n_swap_host, double* n_swap_try_host,
                                 int64_t* n_bad_cov) {
    if (!s) return fail(PV_E_ARG, "pv_sampler_result: sampler is NULL");
    const pv_sampler_dev& d = s->d;
    CUDA_OK(cudaSetDevice(s->p->device));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    const int64_t QC = d.Q * d.C;
    const size_t rows = (size_t)(s->it_host + 1);
    auto down = [&](void* dst, const void* src, size_t bytes) { return !dst || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess; };
    unsigned long long bad = 0;
    long long it_dev = -1;
    bool ok = down(trace_x_host, d.trace_x, sizeof(double) * rows * QC * d.dim) && down(trace_neglogpost_host, d.trace_lp, sizeof(double) * rows * QC) &&
              down(trace_neglogprior_host, d.trace_pr, sizeof(double) * rows * QC) && down(betas_host, d.betas, sizeof(double) * QC) &&
              down(n_acc_host, d.n_acc, sizeof(double) * QC) && down(n_swap_host, d.n_swap, sizeof(double) * d.Q * d.S) &&
              down(n_swap_try_host, d.n_swap_try, sizeof(double) * d.Q * d.S) && down(&bad, d.n_bad_cov, sizeof(bad)) &&
              down(&it_dev, d.it, sizeof(it_dev));
    if (!ok) return fail(PV_E_CUDA, std::string("pv_sampler_result: ") + cudaGetErrorString(cudaGetLastError()));
    if (it_dev != s->it_host + 1) return fail(PV_E_STATE, "pv_sampler_result: device iteration counter out of step with the host");
    if (n_bad_cov) *n_bad_cov = (int64_t)bad;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// device-resident multistart optimiser (pv_optimizer_device.cuh)
// ---------------------------------------------------------------------------------------------------------
struct pv_optimizer {
    pv_problem* p = nullptr;
    pv_opt_dev d;
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    double *d_logp = nullptr, *d_grad = nullptr, *d_fim = nullptr;
    int32_t* d_status = nullptr;
    int iterations = 0;
};

template <class T>
static T* optimizer_alloc(pv_optimizer* o, size_t n) {
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, sizeof(T) * std::max<size_t>(n, 1)) != cudaSuccess) ptr = nullptr;
    if (ptr) cudaMemset(ptr, 0, sizeof(T) * std::max<size_t>(n, 1));
    o->allocs.push_back(ptr);
    return (T*)ptr;
}

static int optimizer_enqueue(pv_optimizer* o, int init) {
    const pv_opt_dev& d = o->d;
    const int64_t QS = d.Q * d.S, B = QS * d.G;
    const unsigned blocks = (unsigned)((QS + 127) / 128);
    pv_opt_propose_kernel<<<blocks, 128, 0, o->stream>>>(d, init);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    if (int rc = launch(o->p, PV_MODE_LOGP_GRAD_FIM, B, d.P, nullptr, nullptr, o->d_logp, o->d_grad, o->d_status, nullptr, o->stream, o->d_fim)) return rc;
    CUDA_OK(cudaMemsetAsync(d.n_active, 0, sizeof(int), o->stream));
    pv_opt_update_kernel<<<blocks, 128, 0, o->stream>>>(d, init);
    g_launches.fetch_add(1);
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" void pv_optimizer_destroy(pv_optimizer* o) {
    if (!o) return;
    if (o->p) cudaSetDevice(o->p->device);
    if (o->stream) {
        cudaStreamSynchronize(o->stream);
        cudaStreamDestroy(o->stream);
    }
    for (void* a : o->allocs) cudaFree(a);
    delete o;
}

extern "C" pv_optimizer* pv_optimizer_create(pv_problem* p, int64_t Q, int n_groups, int n_starts, const double* lb, const double* ub,
                                             const double* prior_loc, const double* prior_scale, const double* x0, double fatol,
                                             double frtol, double gtol) {
    auto bad = [](const char* m) -> pv_optimizer* {
        fail(PV_E_ARG, std::string("pv_optimizer_create: ") + m);
        return nullptr;
    };
    if (!p || !lb || !ub || !prior_loc || !prior_scale || !x0) return bad("NULL argument");
    const int NP = (int)p->col_target.size(), G = n_groups, S = n_starts;
    if (Q < 1 || G < 1 || S < 1) return bad("Q, n_groups, n_starts must be >= 1");
    const int dim = NP * G;
    if (NP < 1 || NP > 8 || dim > PV_OPT_MAXD) return bad("configure the per-group parameters as columns first (n_cols <= 8, n_cols * n_groups <= 16)");
    if (p->stats.empty() || p->n_data != Q * G) return bad("the problem must hold one data set per (problem, group): n_data == Q * n_groups");
    if (cudaSetDevice(p->device) != cudaSuccess) return bad("cudaSetDevice failed");
    auto* o = new pv_optimizer();
    o->p = p;
    memset(&o->d, 0, sizeof(o->d));
    pv_opt_dev& d = o->d;
    d.Q = Q; d.S = S; d.G = G; d.NP = NP; d.dim = dim; d.NS = p->tab->n_sens;
    d.fatol = fatol; d.frtol = frtol; d.gtol = gtol;
    for (int j = 0; j < NP; ++j) {
        int row = -1;
        for (int k = 0; k < p->tab->n_sens; ++k)
            if (p->tab->sens_index[k] == p->col_target[j]) row = k;
        if (row < 0) {
            delete o;
            fail(PV_E_MODEL, "pv_optimizer_create: a column is not a compiled sensitivity parameter of the model");
            return nullptr;
        }
        d.sens_row[j] = row;
    }
    const int64_t QS = Q * S, B = QS * G;
    const int ns = d.NS;
    bool ok = cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking) == cudaSuccess;
    d.x = optimizer_alloc<double>(o, QS * dim); d.f = optimizer_alloc<double>(o, QS); d.g = optimizer_alloc<double>(o, QS * dim);
    d.H = optimizer_alloc<double>(o, QS * dim * dim); d.lam = optimizer_alloc<double>(o, QS);
    d.active = optimizer_alloc<unsigned char>(o, QS); d.converged = optimizer_alloc<unsigned char>(o, QS);
    d.n_iter = optimizer_alloc<int>(o, QS);
    d.x_try = optimizer_alloc<double>(o, QS * dim); d.p = optimizer_alloc<double>(o, QS * dim); d.pred = optimizer_alloc<double>(o, QS);
    d.P = optimizer_alloc<double>(o, (size_t)NP * B);
    o->d_logp = optimizer_alloc<double>(o, B); o->d_grad = optimizer_alloc<double>(o, (size_t)ns * B);
    o->d_fim = optimizer_alloc<double>(o, (size_t)(ns * (ns + 1) / 2) * B); o->d_status = optimizer_alloc<int32_t>(o, B);
    double* d_lb = optimizer_alloc<double>(o, dim); double* d_ub = optimizer_alloc<double>(o, dim);
    double* d_pl = optimizer_alloc<double>(o, Q * dim); double* d_ps = optimizer_alloc<double>(o, Q * dim);
    d.n_evals = optimizer_alloc<unsigned long long>(o, 1); d.n_active = optimizer_alloc<int>(o, 1);
    for (void* a : o->allocs) ok = ok && a != nullptr;
    if (!ok) {
        fail(PV_E_CUDA, std::string("pv_optimizer_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        pv_optimizer_destroy(o);
        return nullptr;
    }
    d.logp = o->d_logp; d.grad = o->d_grad; d.fim = o->d_fim;
    d.lb = d_lb; d.ub = d_ub; d.prior_loc = d_pl; d.prior_scale = d_ps;
    std::vector<double> hx((size_t)QS * dim);
    for (int64_t k = 0; k < QS; ++k)
        for (int i = 0; i < dim; ++i) hx[k * dim + i] = std::min(std::max(x0[k * dim + i], lb[i]), ub[i]);
    auto up = [&](double* dst, const double* src, size_t n) { return cudaMemcpy(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice) == cudaSuccess; };
    ok = up(d.x, hx.data(), hx.size()) && up(d_lb, lb, dim) && up(d_ub, ub, dim) && up(d_pl, prior_loc, (size_t)Q * dim) &&
         up(d_ps, prior_scale, (size_t)Q * dim) && cudaStreamSynchronize(0) == cudaSuccess;
    int rc = ok ? optimizer_enqueue(o, 1) : PV_E_CUDA;          // objective at the start points
    if (!rc && cudaStreamSynchronize(o->stream) != cudaSuccess) rc = PV_E_CUDA;
    if (rc) {
        if (rc == PV_E_CUDA) fail(PV_E_CUDA, std::string("pv_optimizer_create: ") + cudaGetErrorString(cudaGetLastError()));
        pv_optimizer_destroy(o);
        return nullptr;
    }
    return o;
}

extern "C" int pv_optimizer_run(pv_optimizer* o, int max_iter, int poll_every, int* iterations_done) {
    if (!o || max_iter < 0 || poll_every < 1) return fail(PV_E_ARG, "pv_optimizer_run: bad arguments");
    CUDA_OK(cudaSetDevice(o->p->device));
    int n_active = 0;
    CUDA_OK(cudaMemcpyAsync(&n_active, o->d.n_active, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
    CUDA_OK(cudaStreamSynchronize(o->stream));
    int it = 0;
    while (it < max_iter && n_active > 0) {
        const int burst = std::min(poll_every, max_iter - it);
        for (int k = 0; k < burst; ++k)
            if (int rc = optimizer_enqueue(o, 0)) return rc;
        it += burst;
        // iterations past the point where the last start finished change nothing (every thread is masked by `active`)
        CUDA_OK(cudaMemcpyAsync(&n_active, o->d.n_active, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
        CUDA_OK(cudaStreamSynchronize(o->stream));
    }
    o->iterations += it;
    if (iterations_done) *iterations_done = o->iterations;
    return n_active;
}

extern "C" int pv_optimizer_result(pv_optimizer* o, double* x_host, double* f_host, double* g_host, int32_t* n_iter_host,
                                   uint8_t* converged_host, int64_t* n_evals) {
    if (!o) return fail(PV_E_ARG, "pv_optimizer_result: optimizer is NULL");
    const pv_opt_dev& d = o->d;
    CUDA_OK(cudaSetDevice(o->p->device));
    CUDA_OK(cudaStreamSynchronize(o->stream));
    const int64_t QS = d.Q * d.S;
    unsigned long long ne = 0;
    auto down = [&](void* dst, const void* src, size_t bytes) { return !dst || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess; };
    const bool ok = down(x_host, d.x, sizeof(double) * QS * d.dim) && down(f_host, d.f, sizeof(double) * QS) &&
                    down(g_host, d.g, sizeof(double) * QS * d.dim) && down(n_iter_host, d.n_iter, sizeof(int) * QS) &&
                    down(converged_host, d.converged, QS) && down(&ne, d.n_evals, sizeof(ne));
    if (!ok) return fail(PV_E_CUDA, std::string("pv_optimizer_result: ") + cudaGetErrorString(cudaGetLastError()));
    if (n_evals) *n_evals = (int64_t)ne;
    return 0;
}

extern "C" double pv_last_kernel_ms(const pv_problem* p) { return p ? p->last_kernel_ms : -1.0; }

extern "C" double pv_fma_peak_tflops(int device, int use_fp64, int iters) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        fail(PV_E_CUDA, "no CUDA device");
        return -1.0;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    void* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (use_fp64)
            pv_fma_peak_kernel<double><<<blocks, threads>>>((double*)out, iters, 0.999999, 1e-7);
        else
            pv_fma_peak_kernel<float><<<blocks, threads>>>((float*)out, iters, 0.999999f, 1e-7f);
        g_launches.fetch_add(1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}
