/* parvar_b200 - C ABI of the B200-native hot path of matthiaskoenig/parameter-variability.
 *
 * The reference has no FFI of its own for this path: its arithmetic sits behind two Python call
 * shapes that drop into third-party native libraries:
 *
 *   B1  ODESampleSimulator.simulate_samples(parameters, settings)      -> libroadrunner (CVODE)
 *       reference src/parvar/experiments/petab_factory.py:221-272  (hot loop :228-246)
 *   B2  pypesto_problem.objective(x, sensi_orders)                     -> AMICI (CVODES + fwd sens)
 *       reference src/parvar/optimization/petab_optimization.py:43-45 (built), :144-151, :168-173 (called)
 *
 * Each entry point below names the reference call it replaces.  Conventions:
 *   - plain pointers and sizes, no torch / C++ types; the library is loaded with ctypes;
 *   - every function returns int: 0 = ok, < 0 = argument / CUDA error (text via pv_last_error()),
 *     > 0 (only the *_sync helpers) = number of trajectories whose integration failed;
 *   - "device" pointers are CUDA device memory owned by the caller, fp64 unless stated.  Inputs and the
 *     per-trajectory scalars are batch-minor (P[n_cols][B], logp[B], grad[n_sens][B]: consecutive trajectories are
 *     consecutive addresses); trajectories are trajectory-major (Y[B][n_t][n_obs]: one contiguous block per
 *     parameter row, the reference's own sim x time layout, petab_factory.py:269-271);
 *   - launches go to the given stream (a cudaStream_t passed as void*; NULL = legacy default stream),
 *     nothing synchronises unless the name says so;
 *   - a pv_problem may be used from one host thread at a time.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with PV_E_CUDA.
 */
#ifndef PARVAR_B200_H
#define PARVAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_ABI_VERSION 1

enum {
    PV_E_ARG = -1,      /* bad argument */
    PV_E_CUDA = -2,     /* CUDA runtime error */
    PV_E_MODEL = -3,    /* unknown model / id */
    PV_E_STATE = -4,    /* problem not fully configured */
};

/* per-trajectory status codes written to status[B] */
enum { PV_TRAJ_OK = 0, PV_TRAJ_MAX_STEPS = 1, PV_TRAJ_STEP_TOO_SMALL = 2, PV_TRAJ_NONFINITE = 3, PV_TRAJ_STIFF = 4 };

enum { PV_METHOD_AUTO = 0, PV_METHOD_DOPRI5 = 1, PV_METHOD_RODAS4 = 2 };
enum { PV_NOISE_NORMAL = 0, PV_NOISE_LOGNORMAL = 1 };

typedef struct pv_problem pv_problem;

/* ---- library / model registry (reference: parvar.MODELS, src/parvar/__init__.py:26-30) ---------------- */
int pv_abi_version(void);
const char* pv_last_error(void);
int pv_device_count(void);
int pv_model_count(void);
const char* pv_model_name(int model_index);

/* ---- problem = one loaded model + simulation settings ---------------------------------------------------
 * pv_problem_create replaces `roadrunner.RoadRunner(str(model_path))` + integrator settings
 * (petab_factory.py:183-190) and the AMICI model build behind `PetabImporter.create_problem()`
 * (petab_optimization.py:43-45): the model's RHS / Jacobian / sensitivities are already compiled in. */
pv_problem* pv_problem_create(const char* model_name, int device);
void pv_problem_destroy(pv_problem* p);

/* model description (roadrunner `getIds()`, petab_factory.py:190) */
int pv_problem_dims(const pv_problem* p, int* n_states, int* n_theta, int* n_sens);
const char* pv_problem_state_id(const pv_problem* p, int i);
const char* pv_problem_theta_id(const pv_problem* p, int i);
double pv_problem_theta_default(const pv_problem* p, int i);
int pv_problem_sens_theta_index(const pv_problem* p, int k); /* theta index of sensitivity parameter k */
/* algorithmic flop counts measured by the code generator: what[] = "rhs","rhs_sens","jac","lu","solve",
 * "step_dopri5","step_dopri5_sens","step_rodas4","step_rodas4_sens","dense","dense_sens" */
int pv_problem_flops(const pv_problem* p, const char* what);

/* integrator settings (petab_factory.py:187-189: absolute_tolerance, relative_tolerance) */
int pv_problem_set_solver(pv_problem* p, int method, double rtol, double atol, int max_steps);

/* warp scheduling of the batch kernels: a warp fetches new trajectories for its idle lanes once at least
 * refill_min (1..32) of them are idle; 32 = classic one-trajectory-per-thread behaviour per warp; 0 = the
 * measured per-integrator default (DOPRI5: 6, RODAS4: 32) */
int pv_problem_set_schedule(pv_problem* p, int refill_min);

/* Locality schedule (divergence control): the trajectories of a launch are bucketed on the device by their parameter
 * values (first two columns) so that the lanes of a warp integrate neighbouring points of parameter space and therefore
 * step, cross the output grid and finish together (stable radix sort on the device: the order is a pure function of
 * the inputs).  Only the order in which trajectories are picked up changes; results are bit-identical.  mode 0 = auto (launches of >= 32 768 trajectories with per-trajectory columns), 1 = off, 2 = on. */
int pv_problem_set_locality(pv_problem* p, int mode);

/* Lockstep warps (DOPRI5 kernels): the 32 lanes of a warp - 32 consecutive work items, neighbours in parameter space under
 * the locality schedule - advance with ONE step size chosen from the largest error norm among them (warp-level
 * agreement), so t, h, accept / reject and the grid crossings are warp-uniform: no divergence, no refill.  Every lane
 * still meets its own tolerance; a lane's digits beyond the tolerance now depend on its neighbours (a given batch
 * reproduces itself bit for bit: the work order is deterministic).  mode 0 = auto, 1 = off, 2 = on. */
int pv_problem_set_lockstep(pv_problem* p, int mode);

/* DOPRI5 logp + gradient of SMALL batches: one trajectory per group of n_sens neighbouring lanes of a lockstep warp
 * (lane k integrates the state block plus sensitivity block k; only the error norm crosses lanes), launched as one-warp
 * CTAs so that a single-wave launch (BASELINE config 3: 8 x 1000) puts one warp on every scheduler: such a launch lasts
 * as long as its longest trajectory's dependent chain, and this shortens the chain per step.
 * mode 0 = auto (launches that fit one wave), 1 = off, 2 = on. */
int pv_problem_set_group_lanes(pv_problem* p, int mode);

/* Data layout of the Rosenbrock sensitivity kernels (logp + gradient, + Fisher information) of the large models
 * (augmented system of more than 16 components; of the built-in models: icg_body_flat):
 *   PV_LAYOUT_LANE   one trajectory per lane (all blocks of the augmented system in one thread)
 *   PV_LAYOUT_SPLIT  one warp per block: a CTA of 1 + n_sens warps integrates 32 trajectories; stage vectors,
 *                    coefficients and the LU live in shared memory
 *   PV_LAYOUT_AUTO   (default) SPLIT wherever that kernel exists (measured faster at every batch size: 3.6 vs 11 ms
 *                    at 8 000 trajectories, 49 vs 68 ms at 128 000), LANE otherwise.  Both layouts integrate the
 *                    same method with the same tolerances; their results agree within the integration tolerance, not
 *                    bit for bit (different summation order of the stage sums and the error norm). */
#define PV_LAYOUT_AUTO 0
#define PV_LAYOUT_LANE 1
#define PV_LAYOUT_SPLIT 2
int pv_problem_set_layout(pv_problem* p, int layout);

/* Tuning knobs of the launch configuration (tools/sweep_schedule*.py, tools/time_e2e.py):
 *   max_ctas_per_sm  > 0 caps the resident CTAs per SM of the persistent grids (0 = what the occupancy calculator allows)
 *   host_chunk_rows  rows per chunk of the pv_simulate_host pipeline (0 = default 32 768; measured optimum on the C2
 *                    workload at 57 GB/s pinned D2H) */
int pv_problem_set_tuning(pv_problem* p, int max_ctas_per_sm, int64_t host_chunk_rows);

/* PV_METHOD_AUTO on a model classified non-stiff integrates with DOPRI5 first; trajectories that use up
 * switch_steps (default 4000) explicit steps - stiff parameter values, e.g. rate constants near the PEtab upper bound
 * 1e8 that a tempered chain or a uniform start point visits - are re-run with RODAS4 in a second launch on the
 * same stream (the reference's CVODE is implicit throughout).  switch_steps >= max_steps disables the second pass. */
int pv_problem_set_stiff_switch(pv_problem* p, int switch_steps);
/* The explicit first pass of PV_METHOD_AUTO does not have to use up its budget to find out: after `after_steps`
 * attempted steps (default 128) DOPRI5 estimates h*|lambda| from its last two stages at every 4th step (Hairer's
 * stiffness test); 15 consecutive estimates beyond the stability bound 3.25 end the explicit attempt (status
 * PV_TRAJ_STIFF internally) and the trajectory goes to the RODAS4 pass.  A sampler's tempered chains roam the PEtab
 * bounds [0, 1e8]; without the test every such evaluation costs switch_steps explicit steps of pure latency.
 * after_steps = INT32_MAX disables it.  An explicitly selected PV_METHOD_DOPRI5 never runs the test. */
int pv_problem_set_stiffness_detection(pv_problem* p, int after_steps);

/* `r.resetAll(); r.setValue(key, value)` for values shared by the whole batch (dosage dict,
 * petab_factory.py:229-233).  id = a theta id, or a state id / "[state id]" for an initial concentration. */
int pv_problem_set_value(pv_problem* p, const char* id, double value);
int pv_problem_reset_values(pv_problem* p);

/* which ids the columns of the per-trajectory matrix P set (`r.setValue(pid, row[pid])`, :236-239) */
int pv_problem_set_columns(pv_problem* p, int n_cols, const char* const* ids);

/* output grid: `simulate(start, end, steps)` -> linspace(start, end, steps+1) (:242-246); any strictly increasing
 * grid.  t0 <= t[0] is the time at which the initial values hold (roadrunner: `start`; PEtab/AMICI: 0). */
int pv_problem_set_timepoints(pv_problem* p, double t0, int n_t, const double* t_host);

/* selected observables "[sid]" (:250-258); ids may carry the brackets */
int pv_problem_set_observables(pv_problem* p, int n_obs, const char* const* ids);

/* ---- B1: batched forward simulation --------------------------------------------------------------------
 * Replaces the loop `for _, row in parameters.iterrows(): ... s = self.r.simulate(...)` (:228-246).
 *   P_dev      [n_cols][B]   per-trajectory values of the configured columns
 *   Y_dev      [B][n_t][n_obs] concentrations of the selected observables (required)
 *   status_dev [B] int32     PV_TRAJ_* (may be NULL)
 *   steps_dev  [2][B] int32  accepted / rejected step counts (may be NULL)                               */
int pv_simulate(pv_problem* p, int64_t B, const double* P_dev, double* Y_dev, int32_t* status_dev,
                int32_t* steps_dev, void* stream);

/* Same call with HOST buffers (pageable or pinned): chunks the batch, overlaps H2D / kernel / D2H on
 * internal streams and returns when Y_host is complete.  logp_host [B] may be NULL; when given (data must
 * be set) the same pass also returns the log-likelihood.  Returns the number of failed trajectories. */
int pv_simulate_host(pv_problem* p, int64_t B, const double* P_host, double* Y_host, double* logp_host,
                     int32_t* status_host, int32_t* steps_host);

/* ---- B2: log-likelihood (and gradient) ---------------------------------------------------------------------
 * Measurements enter as sufficient statistics per (timepoint, observable, data set):
 *   normal:     n, sum y, sum y^2            lognormal:  n, sum log y, sum (log y)^2  (+ const = -sum log y)
 * so one trajectory can be compared with one individual (n = 1) or with a whole group (pooled fit:
 * every measurement row of a group carries simulationConditionId = group, petab_factory.py:404-417).
 *   stats_host [3][n_t][n_obs][n_data]; trajectory b uses data set  b % n_data.
 *   sigma_host [n_obs]  noise sd per observable (`noiseFormula`, :394; the reference writes 1)
 *   logp_b = sum_{t,k} -0.5*( n log(2 pi sigma^2) + (Syy - 2 f Sy + n f^2)/sigma^2 )          (normal)
 * which is minus AMICI's  0.5*sum(log(2 pi sigma^2) + ((y-f)/sigma)^2).                                   */
int pv_problem_set_data(pv_problem* p, int noise_model, int n_data, const double* stats_host,
                        const double* sigma_host);

/* logp only (`objective(x, sensi_orders=(0,))`, the sampler's call, petab_optimization.py:168-173).
 *   logp_dev [B];  seg_len > 0 additionally sums consecutive runs of seg_len trajectories (one chain's
 *   individuals / one x's groups) into logp_seg_dev [B/seg_len] in a fixed order (deterministic).        */
int pv_logp(pv_problem* p, int64_t B, const double* P_dev, double* logp_dev, int64_t seg_len,
            double* logp_seg_dev, int32_t* status_dev, int32_t* steps_dev, void* stream);

/* logp and gradient by forward sensitivities (`objective(x, sensi_orders=(0,1))`, the optimiser's call,
 * petab_optimization.py:144-151).  grad_dev [n_sens][B] = d logp_b / d theta_s of trajectory b.            */
int pv_logp_grad(pv_problem* p, int64_t B, const double* P_dev, double* logp_dev, double* grad_dev,
                 int64_t seg_len, double* logp_seg_dev, int32_t* status_dev, int32_t* steps_dev, void* stream);

/* logp, gradient and the Fisher information matrix FIM = sum_obs (df/dtheta)(df/dtheta)^T / sigma^2 that AMICI hands
 * to fides as Hessian approximation (`FidesOptimizer`, petab_optimization.py:113-116).
 * fim_dev [n_sens (n_sens + 1) / 2][B]: upper triangle, row-major ((0,0),(0,1),..,(1,1),..).                        */
int pv_logp_grad_fim(pv_problem* p, int64_t B, const double* P_dev, double* logp_dev, double* grad_dev,
                     double* fim_dev, int32_t* status_dev, int32_t* steps_dev, void* stream);

/* trajectories AND logp from one integration pass (BASELINE config 2: "trajectories + logp") */
int pv_simulate_logp(pv_problem* p, int64_t B, const double* P_dev, double* Y_dev, double* logp_dev,
                     int32_t* status_dev, int32_t* steps_dev, void* stream);

/* trajectories and their sensitivities (AMICI `rdata.sy`): Y_dev as above, S_dev [B][n_t][n_obs][n_sens] (both required) */
int pv_simulate_sens(pv_problem* p, int64_t B, const double* P_dev, double* Y_dev, double* S_dev,
                     int32_t* status_dev, int32_t* steps_dev, void* stream);

/* host-buffer versions of the two objective calls (sync; return number of failed trajectories) */
int pv_logp_host(pv_problem* p, int64_t B, const double* P_host, double* logp_host, int64_t seg_len,
                 double* logp_seg_host);
int pv_logp_grad_host(pv_problem* p, int64_t B, const double* P_host, double* logp_host, double* grad_host,
                      int64_t seg_len, double* logp_seg_host);
int pv_logp_grad_fim_host(pv_problem* p, int64_t B, const double* P_host, double* logp_host, double* grad_host,
                          double* fim_host, int64_t seg_len, double* logp_seg_host);

/* ---- hierarchical (population) objective, device resident (BASELINE configs 3-4) -----------
 * Every individual i of chain c has its own theta_ci (the configured columns, which must be compiled sensitivity
 * parameters) and its own data set (individual i of this GPU's shard <-> data set i: n_data == n_local);
 * theta_cip ~ LogNormal(mu_cp, omega_cp).  One call = sensitivity kernel over n_chains * n_local trajectories +
 * an epilogue kernel, both on `stream`, nothing synchronises and nothing touches the host:
 *   theta_dev  [n_cols][n_chains * n_local]   chain-major columns (c * n_local + i)
 *   mu_dev, omega_dev [n_chains][n_cols]      log-scale population mean / sd of every chain
 *   dtheta_dev [n_cols][n_chains * n_local]   out: d/d theta_ci of (log-likelihood + population density)
 *   red_dev    [n_chains][1 + 2 n_cols]       out: this GPU's partial (logp, d/d mu, d/d omega) - the send buffer of
 *                                             the ONE all-reduce (sum) per evaluation when a chain's individuals are
 *                                             sharded over GPUs (ncclAllReduce on the same stream; none on one GPU)
 * A failed integration or theta <= 0 makes the chain's logp NaN.  n_local = 0 (a rank without individuals) writes zeros.
 * Replaces, for all chains at once, what a NUTS / HMC step of the north-star configuration evaluates; the reference's
 * own parallel strategy is a process pool over problems (petab_optimization.py:381-391). */
int pv_hier_logp_grad(pv_problem* p, int n_chains, int64_t n_local, const double* theta_dev, const double* mu_dev,
                      const double* omega_dev, double* dtheta_dev, double* red_dev, int32_t* status_dev,
                      int32_t* steps_dev, void* stream);
/* hyper-priors mu_p ~ Normal(m_p, s_p), omega_p ~ HalfNormal(h_p), added in place AFTER the reduction:
 * red_dev becomes (log posterior, d/d mu, d/d omega) per chain.  The three prior arrays [n_cols] are host memory
 * (they travel in the kernel parameters). */
int pv_hier_finish(int n_chains, int n_cols, double* red_dev, const double* mu_dev, const double* omega_dev,
                   const double* mu_prior_mean_host, const double* mu_prior_sd_host, const double* omega_prior_host,
                   void* stream);

/* ---- Hamiltonian Monte Carlo on the hierarchical posterior, device resident ------------------------------------
 * The consumer the north star names for the logp + gradient op ("the MCMC / NUTS sampler calls at every step"): static-length
 * HMC with jitter, dual-averaging step size, one-window diagonal mass adaptation - optimization/hmc.py states the loop in numpy -
 * with position, momentum, gradient, adaptation state and the kept samples in HBM.  Positive quantities are sampled on the
 * log scale: x_c = (log theta_ci [n_par], mu_c [n_par], log omega_c [n_par]), Jacobian included.  One leapfrog step =
 *     pv_hmc_drift -> pv_hier_logp_grad(theta, mu_q, omega -> dtheta, red) [-> caller's all-reduce of red] -> pv_hier_finish
 *     -> pv_hmc_kick
 * and one iteration = pv_hmc_begin, L leapfrog steps, [caller's all-reduce of energy], pv_hmc_accept - all queued on one
 * stream, no host synchronisation.  Every pointer is device memory (fp64 unless noted); arrays of the shape of theta are
 * [n_par][n_chains * n_local] as pv_hier_logp_grad takes them, hyper arrays [n_chains][2 n_par] = (mu, log omega).
 * The reference has no gradient-based sampler (pyPESTO's adaptive Metropolis, petab_optimization.py:158-173). */
typedef struct pv_hmc_state {
    int32_t n_chains, n_par, own_hyper, reserved;   /* own_hyper: 1 on the ONE rank that accounts for the hyper-parameters' energy terms */
    int64_t n_local, n_total, first;                /* individuals of a chain on this rank / over all ranks / this rank's first */
    double *lth, *mu, *lom, *g_lth, *g_hyp, *lp;    /* state of the chains: position, gradient of the log density, log density [n_chains] */
    double *lth_q, *mu_q, *lom_q, *g_lth_q, *g_hyp_q;   /* trial point of the running trajectory */
    double *p_lth, *p_hyp;                          /* momenta */
    double *im_lth, *im_hyp;                        /* diagonal inverse mass */
    double *eps;                                    /* [n_chains] step size */
    double *theta, *omega, *dtheta, *red;           /* objective buffers: inputs exp(lth_q), exp(lom_q) [n_chains][n_par]; outputs */
    double *energy;                                 /* [n_chains][4]: kinetic energy at start / end, Jacobian term, spare (this rank's partials) */
    int32_t *ok;                                    /* [n_chains] 1 while every evaluation of the trajectory was finite */
    double *da;                                     /* [n_chains][4] dual averaging: mu, log_eps_bar, h_bar, spare */
    double *w_mean_lth, *w_m2_lth, *w_mean_hyp, *w_m2_hyp;   /* Welford moments for the mass matrix */
    double *samples_lth, *samples_hyp, *samples_lp; /* [n_samples][shape of lth] (may be NULL), [n_samples][n_chains][2 n_par] (may be NULL), [n_samples][n_chains] */
    double *acc_sum;                                /* [n_chains] sum of the acceptance probabilities of the kept iterations */
    int32_t *n_div;                                 /* [n_chains] divergent kept iterations */
} pv_hmc_state;
/* momenta = z / sqrt(inverse mass) from this iteration's standard normals z_dev [n_chains][n_total * n_par + 2 n_par] (numpy
 * order of optimization/hmc.py: individual-major theta block, then mu, then log omega), kinetic energy, trial point <- state */
int pv_hmc_begin(const pv_hmc_state* s, const double* z_dev, void* stream);
/* half kick + drift by scale * eps (scale = 0: evaluate the start point), objective inputs */
int pv_hmc_drift(const pv_hmc_state* s, double scale, void* stream);
/* transformed gradient from (dtheta, red), half kick; last != 0: end-of-trajectory energy partials */
int pv_hmc_kick(const pv_hmc_state* s, double scale, int last, void* stream);
/* mode 0: adopt the trial point (start); 1: warm-up iteration it (dual averaging; Welford update number w_n > 0; mass matrix and
 * final step size at it == n_warmup - 1); 2: sampling iteration (kept sample it - n_warmup).  u_dev [n_chains] ~ U(0,1). */
int pv_hmc_accept(const pv_hmc_state* s, int mode, int64_t it, int64_t n_warmup, const double* u_dev, double target_accept,
                  double max_energy_error, int adapt_mass, int64_t w_n, void* stream);

/* ---- No-U-Turn sampler on the hierarchical posterior, device resident ---------------------------------------------
 * BASELINE config 3 names NUTS as the consumer of logp + gradient.  Multinomial NUTS (generalised U-turn criterion, iterative
 * tree building with O(depth) momentum checkpoints, dual averaging, one-window diagonal mass adaptation) as
 * optimization/nuts.py states it in numpy, with every tree vector in HBM and one CTA per chain; all chains advance in lock
 * step, one leapfrog step of every growing chain =
 *     pv_nuts_drift -> pv_hier_logp_grad(theta, mu_q, omega -> dtheta, red) -> pv_hier_finish -> pv_nuts_leaf.
 * The host reads ic[PV_NUTS_GROWING] after a leaf and ic[PV_NUTS_DONE] after a merge (n_chains ints) and draws the uniforms
 * (they travel in the kernel parameters).  Coordinates of a chain, in the order of the flat vectors: log theta_ci
 * (individual-major, n_par each), mu_c [n_par], log omega_c [n_par]; dim = n_ind * n_par + 2 n_par.  One rank (no sharding of
 * the individuals).  No analogue in the reference (pyPESTO's adaptive Metropolis, petab_optimization.py:158-173). */
#define PV_NUTS_MAX_CHAINS 64
enum pv_nuts_vec { PV_NUTS_X, PV_NUTS_G, PV_NUTS_XL, PV_NUTS_RL, PV_NUTS_GL, PV_NUTS_XR, PV_NUTS_RR, PV_NUTS_GR, PV_NUTS_XP,
                   PV_NUTS_GP, PV_NUTS_RSUM, PV_NUTS_XE, PV_NUTS_RE, PV_NUTS_GE, PV_NUTS_SRSUM, PV_NUTS_SX, PV_NUTS_SG,
                   PV_NUTS_XN, PV_NUTS_RN, PV_NUTS_GN, PV_NUTS_IM, PV_NUTS_WMEAN, PV_NUTS_WM2, PV_NUTS_NVEC };
enum pv_nuts_scalar { PV_NUTS_LP, PV_NUTS_H0, PV_NUTS_LPP, PV_NUTS_LOGW, PV_NUTS_SLOGW, PV_NUTS_SLP, PV_NUTS_SUMACC, PV_NUTS_EPS,
                      PV_NUTS_DA_MU, PV_NUTS_DA_LEB, PV_NUTS_DA_HBAR, PV_NUTS_ACCSUM, PV_NUTS_NSC };
enum pv_nuts_int { PV_NUTS_DEPTH, PV_NUTS_DONE, PV_NUTS_DIVERGED, PV_NUTS_NLEAF, PV_NUTS_RIGHT, PV_NUTS_LEAF, PV_NUTS_STURN,
                   PV_NUTS_SDIV, PV_NUTS_GROWING, PV_NUTS_NDIV, PV_NUTS_NIC };
typedef struct pv_nuts_state {
    int32_t n_chains, n_par, max_depth, reserved;
    int64_t n_ind, n_samples;
    double *vec;            /* [PV_NUTS_NVEC][n_chains][dim] tree vectors, state, inverse mass, Welford moments */
    double *ck;             /* [2][max_depth + 1][n_chains][dim] momentum / momentum-sum checkpoints */
    double *sc;             /* [PV_NUTS_NSC][n_chains] */
    int32_t *ic;            /* [PV_NUTS_NIC][n_chains] */
    double *theta, *mu_q, *omega, *dtheta, *red;   /* objective buffers as for pv_hier_logp_grad / pv_hier_finish */
    double *samples;        /* [n_chains][n_samples][dim] */
    double *samples_lp;     /* [n_chains][n_samples] */
    int32_t *samples_depth; /* [n_chains][n_samples] */
    int32_t *host_flags;    /* [2][n_chains] page-locked HOST memory the kernels write directly (or NULL): ic[PV_NUTS_GROWING] after a
                             * leaf, ic[PV_NUTS_DONE] after a merge - the host synchronises the stream and reads them, no copy */
} pv_nuts_state;
typedef struct pv_nuts_uniforms {
    double u[PV_NUTS_MAX_CHAINS];   /* one U(0,1) per chain */
} pv_nuts_uniforms;
/* z_dev [n_chains][dim] standard normals of this iteration */
int pv_nuts_begin(const pv_nuts_state* s, const double* z_dev, void* stream);
int pv_nuts_start(const pv_nuts_state* s, const pv_nuts_uniforms* u_direction, void* stream);
int pv_nuts_drift(const pv_nuts_state* s, void* stream);
int pv_nuts_leaf(const pv_nuts_state* s, const pv_nuts_uniforms* u_take, double max_energy_error, void* stream);
int pv_nuts_merge(const pv_nuts_state* s, const pv_nuts_uniforms* u_swap, void* stream);
/* mode 0: adopt the evaluation of the start point (after drift with every chain not growing, the objective and leaf);
 * 1: warm-up iteration it (Welford update number w_n > 0; mass matrix and final step size at it == n_warmup - 1); 2: sampling */
int pv_nuts_finish(const pv_nuts_state* s, int mode, int64_t it, int64_t n_warmup, double target_accept, int adapt_mass,
                   int64_t w_n, void* stream);

/* page-locked host memory for the *_host calls (cudaHostAlloc / cudaFreeHost); NULL on failure */
void* pv_host_alloc(size_t bytes);
void pv_host_free(void* ptr);

/* ---- host bookkeeping of the batched sampler ------------------------------------------------------------
 * Adaptive Metropolis inside adaptive parallel tempering for Q problems x C temperatures at once; replaces, for all
 * problems together, the per-problem pyPESTO sampler loop behind `pypesto.sample.sample(...)`
 * (src/parvar/optimization/petab_optimization.py:158-173).  Plain host arrays, row-major [Q][C][...]; the objective of
 * the proposals is evaluated by the caller between the two calls (one GPU launch for all of them); random numbers are
 * the caller's (xi ~ N(0,1) [Q][C][dim], u_acc ~ U(0,1) [Q][C], u_swap ~ U(0,1) [Q][C-1]). */
typedef struct pv_sampler_options {
    double decay_constant, target_acceptance_rate, reg_factor, nu, eta;
    int64_t threshold_sample;
    int32_t adapt_betas;
} pv_sampler_options;
/* x_new = x + chol(cov) xi; inside = (lb <= x_new <= ub); x_eval = inside ? x_new : x.  Returns the number of
 * covariance matrices that were not positive definite (their proposal is x), < 0 on argument errors. */
int64_t pv_sampler_propose(int64_t Q, int C, int dim, const double* cov, const double* x, const double* xi,
                           const double* lb, const double* ub, double* x_new, uint8_t* inside, double* x_eval);
/* accept / reject, proposal adaptation (Haario), swaps (hottest pair first), ladder adaptation (Vousden), trace row
 * `it` (1-based, < n_trace) of trace_x [n_trace][Q][C][dim], trace_neglogpost / trace_neglogprior [n_trace][Q][C].
 * In/out: x, llh, lpr, betas, mean_hist, cov_hist, cov_scale, cov, n_acc [Q][C], n_swap / n_swap_try [Q][max(C-1,1)]. */
int pv_sampler_step(int64_t Q, int C, int dim, int64_t it, int64_t n_trace, const pv_sampler_options* options, double* x,
                    double* llh, double* lpr, const double* x_new, const double* llh_new, const double* lpr_new,
                    const uint8_t* inside, const double* u_acc, const double* u_swap, double* betas, double* mean_hist,
                    double* cov_hist, double* cov_scale, double* cov, double* n_acc, double* n_swap, double* n_swap_try,
                    double* trace_x, double* trace_neglogpost, double* trace_neglogprior);

/* ---- the same sampler, device resident --------------------------------------------------------------------
 * Chain state, proposal covariances, ladder and trace live in HBM; one iteration = proposal kernel -> the batched
 * objective of `p` (n_chains * Q * n_groups trajectories, pv_logp) -> accept / adapt / swap / trace kernel, recorded
 * once into a CUDA graph (use_graph != 0) and replayed per iteration on the sampler's own stream; nothing of an
 * iteration touches the host.  Replaces, for Q problems at once, the loop behind `pypesto.sample.sample(...)`
 * (petab_optimization.py:158-173).
 *   p            configured problem: columns = the per-group parameters (n_cols), data = one data set per
 *                (problem, group), n_data == Q * n_groups; x = [n_cols][n_groups] flattened, dim = n_cols * n_groups <= 16
 *   prior_loc / prior_scale [Q][dim]  parameterScaleNormal priors on the lin scale, NaN loc = none (petab_factory.py:443-449)
 *   x0 [Q][dim]  start of every chain of a problem; cov0 [dim][dim] or NULL (identity); ladder: near-exponential, T_max 5e4
 *   rng_block    iterations per uploaded block of random numbers
 * pv_sampler_advance queues n_iter <= rng_block iterations with the caller's random numbers, laid out
 * xi [n_iter][Q][C][dim] ~ N(0,1), u_acc [n_iter][Q][C], u_swap [n_iter][Q][C-1] ~ U(0,1), and returns without waiting for
 * them (it waits for the PREVIOUS block only, to reuse the staging buffers).  Every block but the last must be rng_block long.
 * pv_sampler_result synchronises and copies out rows 0..iterations of the traces ([n_samples+1][Q][C][dim] iteration-major,
 * as pv_sampler_step), the final ladder and the counters; any pointer may be NULL.
 * While a sampler exists its problem must not be used for other launches (they share the problem's scratch buffers). */
typedef struct pv_sampler pv_sampler;
/* numpy's legacy stream (np.random.RandomState) as exported by get_state(): MT19937 key, position, cached deviate */
typedef struct pv_mt19937_state {
    uint32_t key[624];
    int32_t pos, has_gauss;
    double gauss;
} pv_mt19937_state;
pv_sampler* pv_sampler_create(pv_problem* p, int64_t Q, int n_groups, int n_chains, int64_t n_samples,
                              const pv_sampler_options* options, const double* lb, const double* ub, const double* prior_loc,
                              const double* prior_scale, const double* x0, const double* cov0, int64_t rng_block,
                              int use_graph);
int pv_sampler_advance(pv_sampler* s, int64_t n_iter, const double* xi_host, const double* u_acc_host,
                       const double* u_swap_host);
/* pv_sampler_advance with the random numbers drawn by the library itself from numpy's legacy stream (state in / out), in
 * the order optimization/sampler.py draws them, straight into the pinned staging buffers: same chain, no interpreter time.
 * pv_rng_sampler_block is the generator alone (host arrays laid out as pv_sampler_advance takes them). */
int pv_sampler_advance_mt(pv_sampler* s, int64_t n_iter, pv_mt19937_state* rng);
int pv_rng_sampler_block(pv_mt19937_state* rng, int64_t n_iter, int64_t Q, int n_chains, int dim, double* xi_host,
                         double* u_acc_host, double* u_swap_host);
int pv_sampler_result(pv_sampler* s, double* trace_x_host, double* trace_neglogpost_host, double* trace_neglogprior_host,
                      double* betas_host, double* n_acc_host, double* n_swap_host, double* n_swap_try_host,
                      int64_t* n_bad_cov);
void pv_sampler_destroy(pv_sampler* s);

/* ---- batched multistart optimiser, device resident -------------------------------------------------------------
 * Damped Gauss-Newton / Levenberg-Marquardt on the Fisher information + prior precision with box projection (the model
 * fides minimises with FIM Hessian; optimization/optimizer.py states the loop in numpy) for Q problems x n_starts
 * starts at once; replaces `pypesto.optimize.minimize(..., FidesOptimizer, n_starts = 100)` run per problem
 * (petab_optimization.py:99-156).  One iteration = step kernel -> pv_logp_grad_fim over n_starts * Q * n_groups
 * trajectories -> accept / damping / convergence kernel; the host only polls the number of active starts every
 * poll_every iterations.  Problem set-up as for pv_sampler_create (columns must be compiled sensitivity parameters);
 * x0 [Q][n_starts][dim] (clipped to the box), priors [Q][dim] (NaN loc = none).
 * pv_optimizer_run returns the number of starts still active (0 = all converged or stalled), < 0 on error.
 * pv_optimizer_result: x, g [Q][n_starts][dim], f [Q][n_starts] = negative log posterior, in start order (unsorted). */
typedef struct pv_optimizer pv_optimizer;
pv_optimizer* pv_optimizer_create(pv_problem* p, int64_t Q, int n_groups, int n_starts, const double* lb, const double* ub,
                                  const double* prior_loc, const double* prior_scale, const double* x0, double fatol,
                                  double frtol, double gtol);
int pv_optimizer_run(pv_optimizer* o, int max_iter, int poll_every, int* iterations_done);
int pv_optimizer_result(pv_optimizer* o, double* x_host, double* f_host, double* g_host, int32_t* n_iter_host,
                        uint8_t* converged_host, int64_t* n_evals);
void pv_optimizer_destroy(pv_optimizer* o);

/* ---- measurement helpers ------------------------------------------------------------------------------ */
/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t pv_launch_count(void);
/* device time of the kernels of the most recent pv_logp*_host call, in ms (CUDA events); pv_simulate_host pipelines
 * several chunks over three streams and does not update it */
double pv_last_kernel_ms(const pv_problem* p);
/* fp64 / fp32 FMA-pipe peak microbenchmark: returns TFLOP/s achieved by a register-resident FMA chain kernel */
double pv_fma_peak_tflops(int device, int use_fp64, int iters);

#ifdef __cplusplus
}
#endif
#endif /* PARVAR_B200_H */
