/* CPU oracle, C restatement  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may build, load or call this
 * file.  libparvar_b200.so never links it and has no CPU fallback.
 *
 * PARITY UNPINNED (see oracle/parvar_oracle.py for the full statement): the reference's arithmetic
 * for this path lives in libroadrunner 2.9.2 (uv.lock:961-962; SUNDIALS CVODE: variable-order,
 * variable-step BDF with Newton iteration and a dense difference-quotient Jacobian) and in
 * amici 1.0.1 (uv.lock:14-15), neither of which is in /root/reference or installable here, and the
 * reference's tests pin no values.  This file restates
 *   - the three ODE systems exactly as listed in src/parvar/models/sbml/{simple_chain,simple_pk,
 *     icg_body_flat}.md (by hand, independent of the product's SBML reader / code generator);
 *   - the call semantics of ODESampleSimulator.simulate_samples, src/parvar/experiments/
 *     petab_factory.py:221-272: per parameter row resetAll, setValue(dosage), setValue(parameters),
 *     simulate(start, end, steps) on a uniform grid, select observable concentrations;
 *   - the published algorithm class of CVODE's stiff method: a variable-order (1..5) variable-step
 *     BDF in the NDF / backward-difference formulation of Shampine & Reichelt ("The MATLAB ODE
 *     suite", 1997; the same formulation scipy.integrate.BDF uses), simplified Newton with an LU of
 *     I - c J reused across steps, dense difference-quotient Jacobian, local-extrapolation-free
 *     error control in the weighted RMS norm with atol + rtol*|y| weights, polynomial dense output;
 *   - the Gaussian log-likelihood of the PEtab problem the reference writes (petab_factory.py:388-417).
 * It is checked against oracle/parvar_oracle.py (closed forms, scipy LSODA/Radau/BDF) in
 * tests/test_golden.py (C restatement vs the Python oracle and the golden vectors), and is the bench's cpu_baseline ("port", all host cores via OpenMP).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX 12      /* max states */
#define PMAX 40      /* max parameters */
#define MAX_ORDER 5
#define NEWTON_MAXITER 4

typedef void (*rhs_fn)(double t, const double* y, const double* p, double* f);

typedef struct {
    const char* name;
    int nx, np;
    rhs_fn rhs;
    const char* const* states;
    const char* const* params;
    const double* defaults;
    void (*y0)(const double* p, double* y);
    void (*obs_scale)(const double* p, double* s);
} model_t;

/* ---------------------------------------------------------------- simple_chain (simple_chain.md:13-31) */
static const char* const chain_states[] = {"S1", "S2"};
static const char* const chain_params[] = {"k1", "liver"};
static const double chain_defaults[] = {1.0, 1.0};
static void chain_rhs(double t, const double* y, const double* p, double* f) {
    (void)t;
    const double R1 = p[0] * y[0];
    f[0] = -R1 / p[1];
    f[1] = R1 / p[1];
}
static void chain_y0(const double* p, double* y) { (void)p; y[0] = 1.0; y[1] = 0.0; }
static void ones2(const double* p, double* s) { (void)p; s[0] = s[1] = 1.0; }

/* ---------------------------------------------------------------- simple_pk (simple_pk.md:13-41) */
static const char* const pk_states[] = {"y_gut", "y_cent", "y_peri"};
static const char* const pk_params[] = {"CL", "Q", "Vcent", "Vgut", "Vperi", "k_abs"};
static const double pk_defaults[] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
static void pk_rhs(double t, const double* y, const double* p, double* f) {
    (void)t;
    const double CL = p[0], Q = p[1], Vcent = p[2], Vgut = p[3], Vperi = p[4], k_abs = p[5];
    const double ABSORPTION = k_abs * y[0];
    const double CLEARANCE = CL * y[1];
    const double R1 = Q * y[1];
    const double R2 = Q * y[2];
    f[1] = (ABSORPTION / Vcent - CLEARANCE / Vcent - R1 / Vcent) + R2 / Vcent;
    f[0] = -ABSORPTION / Vgut;
    f[2] = R1 / Vperi - R2 / Vperi;
}
/* y_gut: initialAmount = 1 (simple_pk.xml:181) -> concentration 1 / Vgut */
static void pk_y0(const double* p, double* y) { y[0] = 1.0 / p[3]; y[1] = 0.0; y[2] = 0.0; }
static void ones3(const double* p, double* s) { (void)p; s[0] = s[1] = s[2] = 1.0; }

/* ---------------------------------------------------------------- icg_body_flat (icg_body_flat.md:17-141) */
static const char* const icg_states[] = {"Afeces_icg", "Car_icg", "Cgi_plasma_icg", "Chv_icg", "Cli_plasma_icg",
                                         "Clu_plasma_icg", "Cpo_icg", "Cre_plasma_icg", "Cve_icg", "IVDOSE_icg",
                                         "LI__icg", "LI__icg_bi"};
enum { BW, COBW, FQgi, FQh, FQlu, FVar, FVbi, FVgi, FVhv, FVli, FVlu, FVpo, FVve, Fblood, HCT, HEIGHT, ICGIM_Km,
       ICGIM_Vmax, ICGIM_ki_bil, ICGLI2BI_Vmax, ICGLI2CA_Km, ICGLI2CA_Vmax, LI_Vbil, f_oatp1b3, Mr_icg, Ri_icg, Vfeces,
       conv_l_per_ml, conv_ml_per_l, f_bloodflow, f_cardiac_output, f_cirrhosis, f_exercise, f_shunts, f_tissue_loss,
       resection_rate, ti_icg, bil_ext, ICG_NP };
static const char* const icg_params[] = {
    "BW", "COBW", "FQgi", "FQh", "FQlu", "FVar", "FVbi", "FVgi", "FVhv", "FVli", "FVlu", "FVpo", "FVve", "Fblood", "HCT",
    "HEIGHT", "LI__ICGIM_Km", "LI__ICGIM_Vmax", "LI__ICGIM_ki_bil", "LI__ICGLI2BI_Vmax", "LI__ICGLI2CA_Km",
    "LI__ICGLI2CA_Vmax", "LI__Vbil", "LI__f_oatp1b3", "Mr_icg", "Ri_icg", "Vfeces", "conversion_l_per_ml",
    "conversion_ml_per_l", "f_bloodflow", "f_cardiac_output", "f_cirrhosis", "f_exercise", "f_shunts", "f_tissue_loss",
    "resection_rate", "ti_icg", "LI__bil_ext"};
static const double icg_defaults[] = {
    75.0, 1.25, 0.19, 0.255, 1.0, 0.0184, 0.00071, 0.0297, 0.001, 0.021, 0.0076, 0.001, 0.0587, 0.02, 0.51, 170.0,
    0.021659178617926, 0.0369598840327503, 0.02, 0.000114596604507925, 0.012388659243625, 0.000943672769975891, 1.0, 1.0,
    774.96493, 0.0, 1.0, 0.001, 1000.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 5.0, 0.01};

static void icg_rhs(double t, const double* y, const double* p, double* f) {
    (void)t;
    const double Afeces = y[0], Car = y[1], Cgi = y[2], Chv = y[3], Cli = y[4], Clu = y[5], Cpo = y[6], Cre = y[7],
                 Cve = y[8], IVDOSE = y[9], LIicg = y[10], LIicg_bi = y[11];
    (void)Afeces;
    /* parameter-only assignment rules, icg_body_flat.md:79-110 */
    const double CO = p[BW] * p[COBW] * p[f_cardiac_output];
    const double FVre = 1 - (p[FVbi] + p[FVgi] + p[FVli] + p[FVlu] + p[FVve] + p[FVar] + p[FVpo] + p[FVhv]);
    const double Ki_icg = (0.693 / p[ti_icg]) * 60;
    const double fsum = p[FVar] + p[FVve] + p[FVpo] + p[FVhv];
#define VESSEL(FV) ((1 - p[HCT]) * (p[BW] * (FV) - ((FV) / fsum) * p[BW] * p[Fblood] * (1 - fsum)))
    const double Var = VESSEL(p[FVar]);
    const double Vbi = p[BW] * p[FVbi];
    const double Vgi = p[BW] * p[FVgi];
    const double Vhv = VESSEL(p[FVhv]);
    const double Vli = p[BW] * p[FVli] * (1 - p[resection_rate]);
    const double Vlu = p[BW] * p[FVlu];
    const double Vpo = VESSEL(p[FVpo]);
    const double Vve = VESSEL(p[FVve]);
#undef VESSEL
    const double QC = (CO / 1000) * 60;
    const double Vgi_plasma = Vgi * p[Fblood] * (1 - p[HCT]);
    const double Vli_plasma = Vli * p[Fblood] * (1 - p[HCT]);
    const double Vli_tissue = Vli * (1 - p[f_tissue_loss]) * (1 - p[Fblood]);
    const double Vlu_plasma = Vlu * p[Fblood] * (1 - p[HCT]);
    const double Vre = p[BW] * FVre;
    const double Vre_plasma = Vre * p[Fblood] * (1 - p[HCT]);
    const double Qgi = QC * p[FQgi] * p[f_bloodflow];
    const double Qh = QC * p[FQh] * p[f_bloodflow] * p[f_exercise];
    const double Qlu = QC * p[FQlu];
    const double FQre = 1 - Qh / Qlu;
    const double Qpo = Qgi;
    const double Qha = Qh - Qpo;
    const double Qre = QC * FQre;
    /* state-dependent rules, icg_body_flat.md:101-125 */
    const double iv_icg = Ki_icg * IVDOSE / p[Mr_icg];
    const double ICGIM = p[f_oatp1b3] * p[ICGIM_Vmax] * Vli_tissue * Cli /
                         (p[ICGIM_Km] * (1 + p[bil_ext] / p[ICGIM_ki_bil]) + Cli);
    const double ICGLI2BI = p[ICGLI2BI_Vmax] * Vli_tissue * LIicg_bi;
    const double ICGLI2CA = p[ICGLI2CA_Vmax] * Vli_tissue * LIicg / (p[ICGLI2CA_Km] + LIicg);
    const double Flow_ar_argi = Qgi * Car, Flow_gi_po = Qgi * Cgi, Flow_hv_ve = Qh * Chv, Flow_lu_ar = Qlu * Clu,
                 Flow_ve_lu = Qlu * Cve, Flow_po_hv = p[f_shunts] * Qpo * Cpo, Flow_po_li = (1 - p[f_shunts]) * Qpo * Cpo,
                 Flow_ar_arre = Qre * Car, Flow_arli_hv = p[f_shunts] * Qha * Car,
                 Flow_arli_li = (1 - p[f_shunts]) * Qha * Car, Flow_li_hv = (1 - p[f_shunts]) * (Qpo + Qha) * Cli,
                 Flow_re_ve = Qre * Cre;
    /* odes, icg_body_flat.md:127-141 */
    f[0] = ICGLI2BI;
    f[1] = (-Flow_ar_arre / Var - Flow_ar_argi / Var - Flow_arli_li / Var - Flow_arli_hv / Var) + Flow_lu_ar / Var;
    f[2] = Flow_ar_argi / Vgi_plasma - Flow_gi_po / Vgi_plasma;
    f[3] = Flow_arli_hv / Vhv + Flow_po_hv / Vhv + Flow_li_hv / Vhv - Flow_hv_ve / Vhv;
    f[4] = Flow_arli_li / Vli_plasma + Flow_po_li / Vli_plasma - Flow_li_hv / Vli_plasma - ICGIM / Vli_plasma;
    f[5] = Flow_ve_lu / Vlu_plasma - Flow_lu_ar / Vlu_plasma;
    f[6] = Flow_gi_po / Vpo - Flow_po_li / Vpo - Flow_po_hv / Vpo;
    f[7] = Flow_ar_arre / Vre_plasma - Flow_re_ve / Vre_plasma;
    f[8] = iv_icg / Vve + Flow_re_ve / Vve + Flow_hv_ve / Vve - Flow_ve_lu / Vve;
    f[9] = -iv_icg * p[Mr_icg] + p[Ri_icg];
    f[10] = ICGIM / Vli_tissue - ICGLI2CA / Vli_tissue;
    f[11] = ICGLI2CA / Vbi - ICGLI2BI / Vbi;
}
static void icg_y0(const double* p, double* y) { (void)p; for (int i = 0; i < 12; ++i) y[i] = 0.0; }
static void icg_obs(const double* p, double* s) { for (int i = 0; i < 12; ++i) s[i] = 1.0; s[0] = 1.0 / p[Vfeces]; }

static const model_t MODELS[] = {
    {"simple_chain", 2, 2, chain_rhs, chain_states, chain_params, chain_defaults, chain_y0, ones2},
    {"simple_pk", 3, 6, pk_rhs, pk_states, pk_params, pk_defaults, pk_y0, ones3},
    {"icg_body_flat", 12, ICG_NP, icg_rhs, icg_states, icg_params, icg_defaults, icg_y0, icg_obs},
};
#define N_MODELS 3

/* ---------------------------------------------------------------- registry access for the ctypes wrapper */
int pvo_model_index(const char* name) {
    for (int i = 0; i < N_MODELS; ++i)
        if (!strcmp(MODELS[i].name, name)) return i;
    return -1;
}
int pvo_n_states(int m) { return MODELS[m].nx; }
int pvo_n_params(int m) { return MODELS[m].np; }
const char* pvo_state_name(int m, int i) { return MODELS[m].states[i]; }
const char* pvo_param_name(int m, int i) { return MODELS[m].params[i]; }
double pvo_param_default(int m, int i) { return MODELS[m].defaults[i]; }
int pvo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void pvo_rhs(int m, const double* y, const double* p, double* f) { MODELS[m].rhs(0.0, y, p, f); }

/* ---------------------------------------------------------------- dense LU with partial pivoting (n <= NMAX) */
static int lu_factor(int n, double* A, int* piv) {
    for (int k = 0; k < n; ++k) {
        int pk = k;
        double mx = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > mx) { mx = fabs(A[i * n + k]); pk = i; }
        piv[k] = pk;
        if (mx == 0.0) return -1;
        if (pk != k)
            for (int j = 0; j < n; ++j) { double tmp = A[k * n + j]; A[k * n + j] = A[pk * n + j]; A[pk * n + j] = tmp; }
        for (int i = k + 1; i < n; ++i) {
            const double l = A[i * n + k] / A[k * n + k];
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        }
    }
    return 0;
}
static void lu_solve(int n, const double* A, const int* piv, double* b) {
    for (int k = 0; k < n; ++k) {
        const double tmp = b[k]; b[k] = b[piv[k]]; b[piv[k]] = tmp;
        for (int i = k + 1; i < n; ++i) b[i] -= A[i * n + k] * b[k];
    }
    for (int i = n - 1; i >= 0; --i) {
        for (int j = i + 1; j < n; ++j) b[i] -= A[i * n + j] * b[j];
        b[i] /= A[i * n + i];
    }
}

/* ---------------------------------------------------------------- variable-order BDF (NDF), orders 1..5 */
typedef struct {
    int n, order, n_equal_steps, have_lu, current_jac;
    long nfev, njev, nlu, nsteps, nrej;
    double t, h_abs, rtol, atol, newton_tol, t_bound;
    double D[MAX_ORDER + 3][NMAX];
    double J[NMAX * NMAX], LU[NMAX * NMAX];
    int piv[NMAX];
    /* data of the last accepted step for dense output */
    double t_old, h_last;
    int order_last;
    double D_last[MAX_ORDER + 1][NMAX];
    rhs_fn rhs;
    const double* p;
} bdf_t;

static const double KAPPA[MAX_ORDER + 1] = {0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0};
static double GAMMA_[MAX_ORDER + 1], ALPHA[MAX_ORDER + 1], ERRC[MAX_ORDER + 2];
static void bdf_tables(void) {
    GAMMA_[0] = 0;
    for (int k = 1; k <= MAX_ORDER; ++k) GAMMA_[k] = GAMMA_[k - 1] + 1.0 / k;
    for (int k = 0; k <= MAX_ORDER; ++k) {
        ALPHA[k] = (1 - KAPPA[k]) * GAMMA_[k];
        ERRC[k] = KAPPA[k] * GAMMA_[k] + 1.0 / (k + 1);
    }
    ERRC[MAX_ORDER + 1] = 1.0 / (MAX_ORDER + 2);
}

static double rms_norm(int n, const double* x, const double* scale) {
    double s = 0;
    for (int i = 0; i < n; ++i) { const double r = x[i] / scale[i]; s += r * r; }
    return sqrt(s / n);
}

static void compute_R(int order, double factor, double R[MAX_ORDER + 1][MAX_ORDER + 1]) {
    /* M[0,:] = 1; M[i,j] = (i - 1 - factor*j)/i for i,j >= 1; R = cumprod(M, axis 0) */
    for (int j = 0; j <= order; ++j) R[0][j] = 1.0;
    for (int i = 1; i <= order; ++i) {
        R[i][0] = 0.0;
        for (int j = 1; j <= order; ++j) R[i][j] = R[i - 1][j] * ((i - 1 - factor * j) / i);
    }
    /* column 0 of M is zero below row 0 except M[0][0] = 1 -> cumprod gives 1, 0, 0, ... */
}

static void change_D(bdf_t* s, double factor) {
    const int q = s->order, n = s->n;
    double R[MAX_ORDER + 1][MAX_ORDER + 1], U[MAX_ORDER + 1][MAX_ORDER + 1], RU[MAX_ORDER + 1][MAX_ORDER + 1];
    compute_R(q, factor, R);
    compute_R(q, 1.0, U);
    for (int i = 0; i <= q; ++i)
        for (int j = 0; j <= q; ++j) {
            double a = 0;
            for (int k = 0; k <= q; ++k) a += R[i][k] * U[k][j];
            RU[i][j] = a;
        }
    double Dn[MAX_ORDER + 1][NMAX];
    for (int i = 0; i <= q; ++i)
        for (int c = 0; c < n; ++c) {
            double a = 0;
            for (int k = 0; k <= q; ++k) a += RU[k][i] * s->D[k][c];
            Dn[i][c] = a;
        }
    for (int i = 0; i <= q; ++i) memcpy(s->D[i], Dn[i], sizeof(double) * n);
}

/* dense difference-quotient Jacobian, as CVODE's default for dense problems */
static void num_jac(bdf_t* s, double t, const double* y, const double* f0) {
    const int n = s->n;
    double yp[NMAX], f1[NMAX];
    memcpy(yp, y, sizeof(double) * n);
    for (int j = 0; j < n; ++j) {
        const double dy = 1.4901161193847656e-08 * fmax(fabs(y[j]), s->atol / fmax(s->rtol, 1e-300) * 1e-3 + 1e-12);
        yp[j] = y[j] + dy;
        s->rhs(t, yp, s->p, f1);
        s->nfev++;
        for (int i = 0; i < n; ++i) s->J[i * n + j] = (f1[i] - f0[i]) / dy;
        yp[j] = y[j];
    }
    s->njev++;
}

static int bdf_init(bdf_t* s, int n, rhs_fn rhs, const double* p, double t0, const double* y0, double t_bound, double rtol,
                    double atol) {
    memset(s, 0, sizeof(*s));
    s->n = n; s->rhs = rhs; s->p = p; s->t = t0; s->t_bound = t_bound; s->rtol = rtol; s->atol = atol;
    s->newton_tol = fmax(10 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
    double f0[NMAX], scale[NMAX], y1[NMAX], f1[NMAX];
    rhs(t0, y0, p, f0);
    s->nfev++;
    /* initial step (Hairer, Norsett, Wanner I, II.4) with order 1 */
    for (int i = 0; i < n; ++i) scale[i] = atol + fabs(y0[i]) * rtol;
    const double d0 = rms_norm(n, y0, scale), d1 = rms_norm(n, f0, scale);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = fmin(h0, fabs(t_bound - t0));
    for (int i = 0; i < n; ++i) y1[i] = y0[i] + h0 * f0[i];
    rhs(t0 + h0, y1, p, f1);
    s->nfev++;
    double df[NMAX];
    for (int i = 0; i < n; ++i) df[i] = f1[i] - f0[i];
    const double d2 = rms_norm(n, df, scale) / h0;
    const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 0.5);
    s->h_abs = fmin(fmin(100 * h0, h1), fabs(t_bound - t0));
    s->order = 1;
    memcpy(s->D[0], y0, sizeof(double) * n);
    for (int i = 0; i < n; ++i) s->D[1][i] = f0[i] * s->h_abs;
    num_jac(s, t0, y0, f0);
    s->current_jac = 1;
    s->have_lu = 0;
    return 0;
}

/* one accepted step; returns 0 ok, <0 failure */
static int bdf_step(bdf_t* s) {
    const int n = s->n;
    const double min_step = 10 * fabs(nextafter(s->t, INFINITY) - s->t);
    if (s->h_abs < min_step) { s->h_abs = min_step; change_D(s, 1.0); s->n_equal_steps = 0; }
    double y_new[NMAX], d[NMAX], scale[NMAX], y_predict[NMAX], psi[NMAX];
    double t_new = 0, safety = 0, error_norm = 0;
    int n_iter = 0;
    int accepted = 0;
    while (!accepted) {
        if (s->h_abs < min_step) return -1;
        double h = s->h_abs;
        t_new = s->t + h;
        if (t_new - s->t_bound > 0) {
            t_new = s->t_bound;
            change_D(s, fabs(t_new - s->t) / s->h_abs);
            s->n_equal_steps = 0;
            s->have_lu = 0;
        }
        h = t_new - s->t;
        s->h_abs = fabs(h);
        const int q = s->order;
        for (int i = 0; i < n; ++i) {
            double a = 0, b = 0;
            for (int k = 0; k <= q; ++k) a += s->D[k][i];
            for (int k = 1; k <= q; ++k) b += s->D[k][i] * GAMMA_[k];
            y_predict[i] = a;
            psi[i] = b / ALPHA[q];
            scale[i] = s->atol + s->rtol * fabs(a);
        }
        int converged = 0;
        const double c = h / ALPHA[q];
        while (!converged) {
            if (!s->have_lu) {
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) s->LU[i * n + j] = (i == j ? 1.0 : 0.0) - c * s->J[i * n + j];
                if (lu_factor(n, s->LU, s->piv)) return -2;
                s->nlu++;
                s->have_lu = 1;
            }
            /* simplified Newton on  y - y_predict = d,  d = c f(t_new, y) - psi  */
            double y[NMAX], f[NMAX], dy[NMAX];
            memcpy(y, y_predict, sizeof(double) * n);
            memset(d, 0, sizeof(double) * n);
            double dy_norm_old = -1;
            converged = 0;
            int k;
            for (k = 0; k < NEWTON_MAXITER; ++k) {
                s->rhs(t_new, y, s->p, f);
                s->nfev++;
                int finite = 1;
                for (int i = 0; i < n; ++i) finite = finite && isfinite(f[i]);
                if (!finite) break;
                for (int i = 0; i < n; ++i) dy[i] = c * f[i] - psi[i] - d[i];
                lu_solve(n, s->LU, s->piv, dy);
                const double dy_norm = rms_norm(n, dy, scale);
                const double rate = dy_norm_old > 0 ? dy_norm / dy_norm_old : -1;
                if (rate >= 0 && (rate >= 1 || pow(rate, NEWTON_MAXITER - k) / (1 - rate) * dy_norm > s->newton_tol)) break;
                for (int i = 0; i < n; ++i) { y[i] += dy[i]; d[i] += dy[i]; }
                if (dy_norm == 0 || (rate >= 0 && rate / (1 - rate) * dy_norm < s->newton_tol)) { converged = 1; break; }
                dy_norm_old = dy_norm;
            }
            n_iter = k + 1;
            memcpy(y_new, y, sizeof(double) * n);
            if (!converged) {
                if (s->current_jac) break;
                double fp[NMAX];
                s->rhs(t_new, y_predict, s->p, fp);
                s->nfev++;
                num_jac(s, t_new, y_predict, fp);
                s->have_lu = 0;
                s->current_jac = 1;
            }
        }
        if (!converged) {
            s->h_abs *= 0.5;
            change_D(s, 0.5);
            s->n_equal_steps = 0;
            s->have_lu = 0;
            s->nrej++;
            continue;
        }
        safety = 0.9 * (2 * NEWTON_MAXITER + 1) / (2 * NEWTON_MAXITER + n_iter);
        double err[NMAX];
        for (int i = 0; i < n; ++i) { scale[i] = s->atol + s->rtol * fabs(y_new[i]); err[i] = ERRC[q] * d[i]; }
        error_norm = rms_norm(n, err, scale);
        if (error_norm > 1) {
            const double factor = fmax(0.2, safety * pow(error_norm, -1.0 / (q + 1)));
            s->h_abs *= factor;
            change_D(s, factor);
            s->n_equal_steps = 0;
            s->have_lu = 0;
            s->nrej++;
        } else {
            accepted = 1;
        }
    }
    const int q = s->order;
    s->n_equal_steps++;
    s->t_old = s->t;
    s->t = t_new;
    for (int i = 0; i < n; ++i) {
        s->D[q + 2][i] = d[i] - s->D[q + 1][i];
        s->D[q + 1][i] = d[i];
    }
    for (int k = q; k >= 0; --k)
        for (int i = 0; i < n; ++i) s->D[k][i] += s->D[k + 1][i];
    s->nsteps++;
    s->current_jac = 0;
    /* snapshot for dense output */
    s->h_last = s->h_abs;
    s->order_last = q;
    for (int k = 0; k <= q; ++k) memcpy(s->D_last[k], s->D[k], sizeof(double) * n);
    if (s->n_equal_steps < q + 1) return 0;
    /* order selection */
    double scale2[NMAX], e[NMAX];
    for (int i = 0; i < n; ++i) scale2[i] = s->atol + s->rtol * fabs(s->D[0][i]);
    double em = INFINITY, ep = INFINITY;
    if (q > 1) {
        for (int i = 0; i < n; ++i) e[i] = ERRC[q - 1] * s->D[q][i];
        em = rms_norm(n, e, scale2);
    }
    if (q < MAX_ORDER) {
        for (int i = 0; i < n; ++i) e[i] = ERRC[q + 1] * s->D[q + 2][i];
        ep = rms_norm(n, e, scale2);
    }
    const double fm = pow(em, -1.0 / q), f0 = pow(error_norm, -1.0 / (q + 1)), fp = pow(ep, -1.0 / (q + 2));
    int delta = 0;
    double best = f0;
    if (fm > best) { best = fm; delta = -1; }
    if (fp > best) { best = fp; delta = 1; }
    s->order += delta;
    const double factor = fmin(10.0, safety * best);
    s->h_abs *= factor;
    change_D(s, factor);
    s->n_equal_steps = 0;
    s->have_lu = 0;
    return 0;
}

/* polynomial dense output inside the last accepted step */
static void bdf_dense(const bdf_t* s, double t, double* y) {
    const int n = s->n, q = s->order_last;
    const double h = s->h_last;
    double pprod = 1.0;
    for (int i = 0; i < n; ++i) y[i] = s->D_last[0][i];
    for (int k = 0; k < q; ++k) {
        pprod *= (t - (s->t - h * k)) / (h * (1 + k));
        for (int i = 0; i < n; ++i) y[i] += s->D_last[k + 1][i] * pprod;
    }
}

/* ---------------------------------------------------------------- one trajectory on a grid */
static int simulate_one(const model_t* M, const double* p, const double* y0_over, int T, const double* tgrid, double rtol,
                        double atol, double* out /* [T][nx] concentrations */, long* stats) {
    const int n = M->nx;
    double y0[NMAX], sc[NMAX];
    M->y0(p, y0);
    M->obs_scale(p, sc);
    for (int i = 0; i < n; ++i)
        if (y0_over && !isnan(y0_over[i])) y0[i] = y0_over[i];
    int j = 0;
    for (; j < T && tgrid[j] <= tgrid[0]; ++j)
        for (int i = 0; i < n; ++i) out[j * n + i] = y0[i] * sc[i];
    if (j >= T) return 0;
    bdf_t s;
    bdf_init(&s, n, M->rhs, p, tgrid[0], y0, tgrid[T - 1], rtol, atol);
    while (j < T) {
        if (s.nsteps + s.nrej > 1000000) return 1;
        const int rc = bdf_step(&s);
        if (rc) return 2;
        while (j < T && tgrid[j] <= s.t) {
            double y[NMAX];
            if (tgrid[j] == s.t) memcpy(y, s.D[0], sizeof(double) * n); else bdf_dense(&s, tgrid[j], y);
            for (int i = 0; i < n; ++i) out[j * n + i] = y[i] * sc[i];
            ++j;
        }
    }
    if (stats) { stats[0] += s.nsteps; stats[1] += s.nrej; stats[2] += s.nfev; }
    return 0;
}

/* ---------------------------------------------------------------- batch driver (the reference's hot loop)
 * P [n_cols][B] are per-row parameter values for the parameter indices par_idx; theta_base [np] holds the
 * batch-wide values (defaults + dosage-style overrides).  Y [T][n_obs][B] (may be NULL); logp [B] (may be NULL)
 * = Gaussian log-likelihood of `data` [T][n_obs] with sd sigma (petab_factory.py:388-417: sigma = 1). */
int pvo_simulate_batch(int m, long B, int n_cols, const int* par_idx, const double* P, const double* theta_base,
                       const double* y0_over, int T, const double* tgrid, int n_obs, const int* obs_idx, double rtol,
                       double atol, int threads, double* Y, const double* data, double sigma, double* logp, long* stats) {
    if (m < 0 || m >= N_MODELS) return -1;
    const model_t* M = &MODELS[m];
    bdf_tables();
    int failed = 0;
    long st0 = 0, st1 = 0, st2 = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : failed, st0, st1, st2)
    for (long b = 0; b < B; ++b) {
        double p[PMAX];
        memcpy(p, theta_base, sizeof(double) * M->np);
        for (int c = 0; c < n_cols; ++c) p[par_idx[c]] = P[(long)c * B + b];
        double* out = (double*)malloc(sizeof(double) * T * M->nx);
        long st[3] = {0, 0, 0};
        const int rc = simulate_one(M, p, y0_over, T, tgrid, rtol, atol, out, st);
        if (rc) failed++;
        st0 += st[0]; st1 += st[1]; st2 += st[2];
        double lp = 0.0;
        for (int j = 0; j < T; ++j)
            for (int k = 0; k < n_obs; ++k) {
                const double f = out[j * M->nx + obs_idx[k]];
                if (Y) Y[((long)j * n_obs + k) * B + b] = rc ? NAN : f;
                if (logp && data) {
                    const double r = (data[j * n_obs + k] - f) / sigma;
                    lp += -0.5 * (log(2 * M_PI * sigma * sigma) + r * r);
                }
            }
        if (logp) logp[b] = rc ? NAN : lp;
        free(out);
    }
    if (stats) { stats[0] = st0; stats[1] = st1; stats[2] = st2; }
    return failed;
}
