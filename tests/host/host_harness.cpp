// Host-side compile of the *device* integrator templates + generated model functions, so their
// numerics can be checked with g++ in the GPU-less build container (tests/test_host_numerics.py).
// Test infrastructure only: never linked into libparvar_b200.so, never used by the product.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../parameter-variability_b200/csrc/pv_dopri5.cuh"
#include "../../parameter-variability_b200/csrc/pv_rosenbrock.cuh"
#include "../../parameter-variability_b200/csrc/generated/simple_chain.cuh"
#include "../../parameter-variability_b200/csrc/generated/simple_pk.cuh"
#include "../../parameter-variability_b200/csrc/generated/icg_body_flat.cuh"

template <int N>
struct Collect {
    double* out;
    void operator()(int j, double, const double (&z)[N]) {
        for (int i = 0; i < N; ++i) out[j * N + i] = z[i];
    }
};

static int g_stiff_after = 0x7fffffff;   // DOPRI5 stiffness detection of the next hh_simulate call (hh_set_stiff_after)
extern "C" void hh_set_stiff_after(int n) { g_stiff_after = n; }

template <class M, bool SENS>
static int run(int method, const double* theta, const double* y0_over, const double* tgrid, int T, double rtol,
               double atol, int max_steps, double* out, int* stats) {
    using Sys = pv_system<M, SENS>;
    double th[M::NTH];
    for (int i = 0; i < M::NTH; ++i) th[i] = theta[i];
    typename Sys::Consts ks;
    double y0[M::NX], obs[M::NX];
    M::hoist(th, ks.c, y0, obs);
    double z[Sys::N];
    for (int i = 0; i < M::NX; ++i) z[i] = std::isnan(y0_over[i]) ? y0[i] : y0_over[i];
    if constexpr (SENS) {
        double dy0[M::NSD][M::NX];
        M::hoist_sens(th, ks.dc, dy0);
        for (int p = 0; p < M::NS; ++p)
            for (int i = 0; i < M::NX; ++i) z[M::NX * (1 + p) + i] = std::isnan(y0_over[i]) ? dy0[p][i] : 0.0;
    }
    Collect<Sys::N> col{out};
    pv_step_stats st;
    int rc;
    if (method == 1) {
        pv_dopri5_lane<Sys> lane;
        lane.stiff_after = g_stiff_after;
        rc = lane.start(ks, z, tgrid[0], tgrid, T, rtol, atol, col);
        while (rc == PV_RUNNING) rc = lane.step(ks, tgrid, T, rtol, atol, max_steps, col);
        st = lane.st;
    } else
        rc = pv_rosenbrock_integrate<M, SENS>(ks, z, tgrid[0], tgrid, T, rtol, atol, max_steps, col, st);
    stats[0] = st.accepted;
    stats[1] = st.rejected;
    // apply observable scaling to the state block (and sensitivity blocks)
    for (int j = 0; j < T; ++j)
        for (int b = 0; b < Sys::N / M::NX; ++b)
            for (int i = 0; i < M::NX; ++i) out[j * Sys::N + b * M::NX + i] *= obs[i];
    return rc;
}

template <class M>
static int info(int* dims, double* /*unused*/) {
    dims[0] = M::NX; dims[1] = M::NTH; dims[2] = M::NS; dims[3] = M::NC; dims[4] = M::NNZ; dims[5] = M::NLU;
    return 0;
}

#define DISPATCH(CALL)                                                     \
    if (!strcmp(model, "simple_chain")) return CALL(pv_model_simple_chain); \
    if (!strcmp(model, "simple_pk")) return CALL(pv_model_simple_pk);       \
    if (!strcmp(model, "icg_body_flat")) return CALL(pv_model_icg_body_flat);

extern "C" int hh_dims(const char* model, int* dims) {
#define C1(M) info<M>(dims, nullptr)
    DISPATCH(C1)
    return -1;
}

extern "C" int hh_simulate(const char* model, int method, int sens, const double* theta, const double* y0_over,
                           const double* tgrid, int T, double rtol, double atol, int max_steps, double* out,
                           int* stats) {
#define C2(M) (sens ? run<M, true>(method, theta, y0_over, tgrid, T, rtol, atol, max_steps, out, stats) \
                    : run<M, false>(method, theta, y0_over, tgrid, T, rtol, atol, max_steps, out, stats))
    DISPATCH(C2)
    return -1;
}

// n fixed RODAS4 steps of size h (step-size control bypassed) -> final state; for the order check
template <class M>
static int fixed_steps(const double* theta, const double* y0_over, double h, int n, double* out) {
    using Sys = pv_system<M, false>;
    double th[M::NTH];
    for (int i = 0; i < M::NTH; ++i) th[i] = theta[i];
    typename Sys::Consts ks;
    double y0[M::NX], obs[M::NX], z[M::NX];
    M::hoist(th, ks.c, y0, obs);
    for (int i = 0; i < M::NX; ++i) z[i] = std::isnan(y0_over[i]) ? y0[i] : y0_over[i];
    double tg[2] = {0.0, h * n};
    std::vector<double> sink_buf(2 * M::NX);
    Collect<M::NX> col{sink_buf.data()};
    pv_rodas4_lane<M, false> lane;
    int rc = lane.start(ks, z, 0.0, tg, 2, 1e30, 1e30, col);
    for (int k = 0; k < n && rc == PV_RUNNING; ++k) {
        lane.h = h;
        rc = lane.step(ks, tg, 2, 1e30, 1e30, 1 << 30, col);
    }
    for (int i = 0; i < M::NX; ++i) out[i] = lane.z[i] * obs[i];
    return rc;
}

extern "C" int hh_rodas4_fixed(const char* model, const double* theta, const double* y0_over, double h, int n, double* out) {
#define C4(M) fixed_steps<M>(theta, y0_over, h, n, out)
    DISPATCH(C4)
    return -1;
}

// raw model functions for codegen checks
extern "C" int hh_rhs(const char* model, const double* theta, const double* y, double* f, double* Jdense) {
#define C3(M) ([&]() {                                                              \
        double th[M::NTH], c[M::NC], y0[M::NX], obs[M::NX], yy[M::NX], ff[M::NX], J[M::NNZ]; \
        for (int i = 0; i < M::NTH; ++i) th[i] = theta[i];                           \
        for (int i = 0; i < M::NX; ++i) yy[i] = y[i];                                \
        M::hoist(th, c, y0, obs);                                                    \
        M::rhs(0.0, yy, c, ff);                                                      \
        M::jac(0.0, yy, c, J);                                                       \
        for (int i = 0; i < M::NX; ++i) f[i] = ff[i];                                \
        /* check the LU: solve (d I - J) x = f with d = 3 and return x in Jdense[0..NX) */ \
        double W[M::NLU], b[M::NX];                                                  \
        M::lu_assemble(3.0, J, W); M::lu_factor(W);                                  \
        for (int i = 0; i < M::NX; ++i) b[i] = ff[i];                                \
        M::lu_solve(W, b);                                                           \
        for (int i = 0; i < M::NX; ++i) Jdense[i] = b[i];                            \
        for (int n = 0; n < M::NNZ; ++n) Jdense[M::NX + n] = J[n];                   \
        return 0; })()
    DISPATCH(C3)
    return -1;
}
