"""Host logic of the SURVEY 8(f) "next" rows (sampler, optimiser, PEtab tables, result helpers) - no GPU needed:
the objective is replaced by analytic functions."""
import numpy as np
import pandas as pd
import pytest

from conftest import pkg


def test_adaptive_metropolis_parallel_tempering_on_gaussian():
    smp = pkg("optimization.sampler")
    mean = np.array([[1.0, -2.0], [0.5, 3.0]])                 # two problems
    sd = np.array([[0.5, 2.0], [1.5, 0.2]])

    def evaluate(X):                                           # X [Q, C, dim]
        z = (X - mean[:, None, :]) / sd[:, None, :]
        return -0.5 * np.sum(z * z, axis=2), np.zeros(X.shape[:2])

    lb, ub = np.full(2, -50.0), np.full(2, 50.0)
    res = smp.sample(evaluate, x0=mean + 0.1, lb=lb, ub=ub, n_samples=6000, n_chains=4, seed=1)
    assert res.trace_x.shape == (2, 4, 6001, 2) and res.n_evals == 2 * 4 * 6001
    cold = res.trace_x[:, 0, 1000:, :]
    np.testing.assert_allclose(cold.mean(axis=1), mean, atol=0.25 * sd.max())
    np.testing.assert_allclose(cold.std(axis=1), sd, rtol=0.25)
    assert np.all(res.acceptance_rate[:, 0] > 0.1) and np.all(res.acceptance_rate[:, 0] < 0.5)
    assert np.all(res.betas[:, 0] == 1.0) and np.all(np.diff(res.betas, axis=1) < 0)
    hot = res.trace_x[:, -1, 1000:, :]
    assert np.all(hot.std(axis=1) > 2 * sd)                   # the hottest chain explores far more widely
    ess = smp.effective_sample_size(res)
    assert np.all(ess > 50) and np.all(ess < 6001)
    res2 = smp.sample(evaluate, x0=mean + 0.1, lb=lb, ub=ub, n_samples=200, n_chains=4, seed=1)
    res3 = smp.sample(evaluate, x0=mean + 0.1, lb=lb, ub=ub, n_samples=200, n_chains=4, seed=1)
    assert np.array_equal(res2.trace_x, res3.trace_x)         # seeded -> reproducible


def test_sampler_respects_bounds_and_failed_evaluations():
    smp = pkg("optimization.sampler")

    def evaluate(X):
        llh = -0.5 * np.sum(X * X, axis=2)
        llh[X[..., 0] > 0.8] = -np.inf                        # "integration failed" region
        return llh, np.zeros(X.shape[:2])

    res = smp.sample(evaluate, x0=np.array([[0.1, 0.1]]), lb=np.array([-1.0, 0.0]), ub=np.array([1.0, 1.0]), n_samples=1500, seed=3)
    x = res.trace_x[0, 0]
    assert x[:, 0].min() >= -1.0 and x[:, 1].min() >= 0.0 and x[:, 1].max() <= 1.0 and x[:, 0].max() <= 0.8


def test_ladder_hdi_and_autocorrelation():
    smp = pkg("optimization.sampler")
    b = smp.near_exponential_decay_betas(4)
    assert b[0] == 1.0 and b[-1] == pytest.approx(1 / 5e4) and np.all(np.diff(b) < 0)
    assert np.array_equal(smp.near_exponential_decay_betas(1), [1.0])
    rs = np.random.RandomState(0)
    lo, hi = smp.hdi(rs.normal(size=200000))
    assert lo == pytest.approx(-1.88, abs=0.05) and hi == pytest.approx(1.88, abs=0.05)
    assert smp.autocorrelation_time_sokal(rs.normal(size=20000)) == pytest.approx(1.0, abs=0.2)
    ar = np.zeros(50000)
    for i in range(1, len(ar)):
        ar[i] = 0.9 * ar[i - 1] + rs.normal()
    assert smp.autocorrelation_time_sokal(ar) == pytest.approx((1 + 0.9) / (1 - 0.9), rel=0.2)


def test_batched_trust_region_optimizer():
    om = pkg("optimization.optimizer")
    t = np.linspace(0, 4, 30)
    truth = np.array([[2.0, 0.7], [1.0, 1.5]])                # two problems: y = a * exp(-k t)

    def evaluate(X):                                           # X [Q, S, 2]
        a, k = X[..., 0:1], X[..., 1:2]
        y = truth[:, None, 0:1] * np.exp(-truth[:, None, 1:2] * t)
        m = a * np.exp(-k * t)
        r = m - y
        J = np.stack([np.exp(-k * t), -a * t * np.exp(-k * t)], axis=-1)        # [Q, S, T, 2]
        return {"fval": 0.5 * np.sum(r * r, axis=-1), "grad": np.einsum("qst,qstd->qsd", r, J),
                "hess": np.einsum("qstd,qste->qsde", J, J)}

    lb, ub = np.array([0.0, 0.0]), np.array([10.0, 10.0])
    x0 = om.uniform_startpoints(16, lb, ub, Q=2, seed=5)
    res = om.minimize(evaluate, x0, lb, ub, maxiter=200)
    np.testing.assert_allclose(res.x[:, 0, :], truth, rtol=1e-6)
    assert np.all(res.fval[:, 0] < 1e-12) and np.all(np.diff(res.fval, axis=1) >= 0)      # sorted, best first
    assert res.converged[:, 0].all()
    # a bound-constrained optimum: true k above the upper bound
    res = om.minimize(evaluate, x0, lb, np.array([10.0, 0.5]), maxiter=200)
    assert np.allclose(res.x[0, 0, 1], 0.5)


def test_petab_tables_roundtrip(tmp_path):
    io = pkg("optimization.petab_io")
    po = pkg("optimization.petab_optimization")
    P = pkg()
    rs = np.random.RandomState(0)
    t = np.linspace(0, 100, 5)
    groups = [po.GroupData("MALE", rs.rand(3, 5, 2)), po.GroupData("FEMALE", rs.rand(3, 5, 2))]
    yaml_path = io.write_petab_like_reference(tmp_path / "abc123", P.MODELS["simple_pk"], groups, t, ["y_cent", "y_gut"],
                                              ["k_abs", "CL"], {"k_abs_MALE": (0.5, 1.0), "CL_FEMALE": (0.5, 1.0)})
    tb = io.read_petab_tables(yaml_path)
    assert len(tb.measurements) == 2 * 3 * 5 * 2
    assert list(tb.conditions["conditionId"]) == ["MALE", "FEMALE"] and list(tb.conditions["k_abs"]) == ["k_abs_MALE", "k_abs_FEMALE"]
    assert set(tb.conditions.columns) >= {"y_cent", "y_gut"} and (tb.conditions["y_cent"] == 0.0).all()      # quirk Q4
    assert list(tb.parameters["parameterId"]) == ["k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE"]
    assert tb.parameters.set_index("parameterId").loc["k_abs_MALE", "objectivePriorParameters"] == "0.5;1.0"
    assert tb.model_path.exists()


def test_result_helpers():
    d = pkg("optimization.driver")
    assert d.get_group_from_pid("k_abs_MALE") == "MALE" and d.get_parameter_from_pid("k_abs_MALE") == "k_abs"
    assert d.get_parameter_from_pid("LI__ICGIM_Vmax_FEMALE") == "LI__ICGIM_Vmax"
    assert d.get_parameter_from_pid("k1") == "k1"
    assert d.DEFAULT_SETTINGS["n_starts"] == 100 and d.DEFAULT_SETTINGS["n_samples"] == 5000 and d.DEFAULT_SETTINGS["n_chains"] == 4


def test_batched_hmc_on_correlated_gaussian():
    hm = pkg("optimization.hmc")
    A = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, -0.3], [0.0, -0.3, 0.5]])
    Pm = np.linalg.inv(A)
    mean = np.array([0.5, -1.0, 2.0])

    def lg(X):
        d = X - mean
        return -0.5 * np.einsum("ci,ij,cj->c", d, Pm, d), -d @ Pm

    res = hm.hmc(lg, x0=np.zeros((6, 3)), n_samples=1500, n_warmup=600, n_leapfrog=12, seed=3)
    s = res.samples.reshape(-1, 3)
    np.testing.assert_allclose(s.mean(axis=0), mean, atol=0.12)
    np.testing.assert_allclose(np.cov(s.T), A, atol=0.25)
    assert np.all(res.accept_rate > 0.6) and np.all(res.accept_rate < 0.98) and res.divergences.sum() == 0
    assert res.n_grad_evals > 1500 * 8
    # failed gradient evaluations (non-finite) count as divergences and are rejected, not propagated
    def lg_bad(X):
        l, g = lg(X)
        l = np.where(X[:, 0] > 1.5, np.nan, l)
        return l, g
    res = hm.hmc(lg_bad, x0=np.zeros((2, 3)), n_samples=300, n_warmup=200, n_leapfrog=8, seed=4)
    assert np.all(np.isfinite(res.samples)) and res.samples[..., 0].max() <= 1.5


def test_batched_nuts_on_correlated_gaussian():
    nt = pkg("optimization.nuts")
    A = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, -0.3], [0.0, -0.3, 0.5]])
    Pm = np.linalg.inv(A)
    mean = np.array([0.5, -1.0, 2.0])
    calls = []

    def lg(X):
        calls.append(X.shape)
        d = X - mean
        return -0.5 * np.einsum("ci,ij,cj->c", d, Pm, d), -d @ Pm

    res = nt.nuts(lg, x0=np.zeros((6, 3)), n_samples=1200, n_warmup=500, seed=3)
    s = res.samples.reshape(-1, 3)
    np.testing.assert_allclose(s.mean(axis=0), mean, atol=0.1)
    np.testing.assert_allclose(np.cov(s.T), A, atol=0.2)
    assert np.all(res.accept_rate > 0.65) and np.all(res.accept_rate < 0.95) and res.divergences.sum() == 0
    assert res.tree_depth.min() >= 1 and res.tree_depth.max() <= 8 and 1.5 < res.tree_depth.mean() < 5
    assert all(c == (6, 3) for c in calls) and res.n_grad_evals == len(calls)      # every evaluation is a full batch
    # independent draws: lag-1 autocorrelation of NUTS on a Gaussian is small
    c0 = res.samples[0, :, 1] - res.samples[0, :, 1].mean()
    assert abs(np.dot(c0[1:], c0[:-1]) / np.dot(c0, c0)) < 0.35
    # heavy ridge: a banana-shaped target exercises deeper trees without divergences
    def banana(X):
        a, b = X[:, 0], X[:, 1]
        lp = -0.5 * (a * a / 4.0 + (b - 0.25 * a * a) ** 2 / 0.25)
        g = np.stack([-(a / 4.0) + (b - 0.25 * a * a) / 0.25 * 0.5 * a, -(b - 0.25 * a * a) / 0.25], axis=1)
        return lp, g
    rb = nt.nuts(banana, x0=np.zeros((4, 2)), n_samples=600, n_warmup=400, seed=9, adapt_mass=False, target_accept=0.9)
    sb = rb.samples.reshape(-1, 2)
    assert abs(sb[:, 0].mean()) < 0.35 and abs(sb[:, 0].var() - 4.0) < 1.2 and rb.divergences.sum() <= 0.03 * sb.shape[0]


def test_reference_analysis_layout(tmp_path):
    """The files the reference's ``join_optimization_results`` (analysis/utils.py:28-98) reads: same paths, same columns;
    the merge it performs (on id, group, parameter) must give one row per (experiment, group, parameter)."""
    import ast
    import json
    d = pkg("optimization.driver")
    xps = [{"id": f"uid{k}", "model": "simple_pk", "prior_type": "exact_prior", "samples": 20, "timepoints": 10 + k, "noise_cv": 0.05,
            "sampling": {"MALE": {"CL": (1.0, 1.0), "k_abs": (0.5, 1.0)}, "FEMALE": {"CL": (0.5, 1.0), "k_abs": (1.0, 1.0)}}} for k in range(3)]
    frames = {}
    for xp in xps:
        rows = []
        for pid in ("k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE"):
            rows.append({"id": xp["id"], "group": d.get_group_from_pid(pid), "parameter": d.get_parameter_from_pid(pid), "pid": pid,
                         "mean": 1.0, "std": 0.1, "median": 1.0, "ess": 100.0, "n_samples": 50, "optim_duration": 0.1,
                         "sampler_duration": 0.2, "values": np.ones((4, 51)), "hdi_low": 0.8, "hdi_high": 1.2})
        frames[xp["id"]] = pd.DataFrame(rows)
    base = d.write_reference_layout(tmp_path, "all", xps, frames)
    # --- what join_optimization_results does -------------------------------------------------------------------------
    df_bayes = pd.concat(pd.read_csv(f, sep="\\t") for f in sorted((base / "optimization_results").glob("*.tsv")))
    df_bayes["values"] = df_bayes["values"].apply(lambda x: np.array(json.loads(x)))
    df_xp = pd.read_csv(base / "definitions.tsv", sep="\\t")
    assert list(df_xp.columns) == ["id", "model", "prior_type", "n_groups", "group", "samples", "timepoints", "noise_cv",
                                   "parameter", "dsn_type", "dsn_par"]
    df_join = df_xp.merge(df_bayes, on=["id", "group", "parameter"], how="inner")
    assert len(df_join) == 3 * 4
    assert df_join["dsn_par"].apply(lambda s: ast.literal_eval(s)["loc"]).notna().all()
    for col in ("optim_duration", "mean", "median", "n_samples", "values", "ess", "hdi_high", "hdi_low"):
        assert col in df_join.columns
    assert df_join["values"].iloc[0].shape == (4, 51)


def test_hierarchical_population_terms_host_algebra():
    """HierarchicalObjective.local_terms (host part, kernel replaced by a stub): the population density of
    theta_ip ~ LogNormal(mu_p, omega_p) and its derivatives against the plain formula and central differences;
    outputs are C-contiguous (they are packed into all-reduce buffers)."""
    po = pkg("optimization.petab_optimization")

    class Stub:                                   # stands in for runtime.Problem: zero likelihood
        columns = ["a", "b"]
        sens_ids = ["b", "a"]                     # kernel's sensitivity order differs from the column order
        n_sens = 2

        def logp_host(self, Pm, seg_len=0, grad=False):
            B = Pm.shape[1]
            self.last_P = Pm.copy()
            return np.zeros(B), np.zeros((2, B)), np.zeros(B // seg_len), 0

    C, L, P = 3, 7, 2
    rs = np.random.RandomState(5)
    theta = rs.lognormal(size=(C, L, P))
    mu, omega = rs.normal(size=(C, P)), rs.uniform(0.5, 2.0, size=(C, P))
    stub = Stub()
    H = po.HierarchicalObjective(stub, L, mu_prior=[(0.0, 0.5)] * P, omega_prior=[0.5] * P)

    def pop(th, m, w):
        z = (np.log(th) - m[:, None, :]) / w[:, None, :]
        return np.sum(-np.log(th) - np.log(w)[:, None, :] - 0.5 * np.log(2 * np.pi) - 0.5 * z * z, axis=(1, 2))

    lp, dth, dmu, dom = H.local_terms(theta, mu, omega)
    for a in (lp, dth, dmu, dom):
        assert a.flags["C_CONTIGUOUS"]
    assert dth.shape == (C, L, P) and dmu.shape == (C, P) and dom.shape == (C, P)
    np.testing.assert_array_equal(stub.last_P, theta.transpose(2, 0, 1).reshape(P, C * L))    # parameter-major rows
    np.testing.assert_allclose(lp, pop(theta, mu, omega), rtol=1e-13)
    eps = 1e-6
    for c, l, p in [(0, 0, 0), (2, 6, 1), (1, 3, 0)]:
        tp, tm = theta.copy(), theta.copy()
        tp[c, l, p] += eps
        tm[c, l, p] -= eps
        np.testing.assert_allclose(dth[c, l, p], (pop(tp, mu, omega)[c] - pop(tm, mu, omega)[c]) / (2 * eps), rtol=1e-6)
    for c, p in [(0, 0), (2, 1)]:
        mp, mm = mu.copy(), mu.copy()
        mp[c, p] += eps
        mm[c, p] -= eps
        np.testing.assert_allclose(dmu[c, p], (pop(theta, mp, omega)[c] - pop(theta, mm, omega)[c]) / (2 * eps), rtol=1e-6)
        wp, wm = omega.copy(), omega.copy()
        wp[c, p] += eps
        wm[c, p] -= eps
        np.testing.assert_allclose(dom[c, p], (pop(theta, mu, wp)[c] - pop(theta, mu, wm)[c]) / (2 * eps), rtol=1e-6)


def test_native_sampler_bookkeeping_matches_numpy():
    """The compiled host loop of the sampler (pv_sampler_propose / pv_sampler_step, include/parvar_b200.h) and its numpy
    statement draw the same random numbers in the same order and must walk the same chains: identical accept / swap
    decisions, positions equal up to round-off - including proposals outside the bounds, failed evaluations (NaN),
    a single temperature, and covariance matrices that need the eigenvalue floor."""
    smp = pkg("optimization.sampler")
    A = np.array([[2.0, 0.6, 0.1], [0.6, 1.0, 0.3], [0.1, 0.3, 0.5]])
    Ai = np.linalg.inv(A)

    def evaluate(X):
        d = X - 0.3
        llh = -0.5 * np.einsum("qci,ij,qcj->qc", d, Ai, d) * 40
        llh = np.where(X[..., 0] > 1.2, np.nan, llh)                      # a region where the "simulation fails"
        return llh, -0.5 * np.sum((X - 0.5) ** 2, axis=2)

    Q, dim = 24, 3
    x0 = np.full((Q, dim), 0.25)
    lb, ub = np.zeros(dim), np.array([2.0, 1e8, 1e8])                     # tight lower bound: many rejected proposals
    for chains in (4, 1):
        r0 = smp.sample(evaluate, x0, lb, ub, n_samples=300, n_chains=chains, native=False)
        r1 = smp.sample(evaluate, x0, lb, ub, n_samples=300, n_chains=chains, native=True)
        np.testing.assert_array_equal(r0.acceptance_rate, r1.acceptance_rate)
        np.testing.assert_array_equal(r0.swap_rate, r1.swap_rate)
        np.testing.assert_allclose(r1.trace_x, r0.trace_x, rtol=0, atol=1e-9)
        np.testing.assert_allclose(r1.trace_neglogpost, r0.trace_neglogpost, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(r1.trace_neglogprior, r0.trace_neglogprior, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(r1.betas, r0.betas, rtol=1e-12)
        assert r1.trace_x.shape == (Q, chains, 301, dim) and r1.trace_x.flags["C_CONTIGUOUS"]
        assert 0.05 < r1.acceptance_rate[:, 0].mean() < 0.9
    # degenerate start covariance (rank 1, scale 1e15): the eigenvalue floor of _regularize / pv_sampler::regularize
    v = np.array([1.0, 1.0, 1.0])
    cov0 = 1e15 * np.outer(v, v)
    ra = smp.sample(evaluate, x0, lb, np.full(dim, 1e8), n_samples=50, cov0=cov0, native=False)
    rb = smp.sample(evaluate, x0, lb, np.full(dim, 1e8), n_samples=50, cov0=cov0, native=True)
    np.testing.assert_array_equal(ra.acceptance_rate, rb.acceptance_rate)
    np.testing.assert_allclose(rb.trace_x, ra.trace_x, rtol=1e-6, atol=1e-6)


def test_covariance_regularisation_and_sampler_abi_errors():
    """_regularize: symmetric output, eigenvalues >= max(reg, 1e-12 lambda_max), good matrices untouched; the batched
    Cholesky test agrees with the eigenvalues.  pv_sampler_propose: proposals outside the box are flagged and replaced
    by the current point in x_eval, a covariance that is not positive definite is counted and proposes x itself;
    argument errors come back as negative codes with a message."""
    smp = pkg("optimization.sampler")
    rt = pkg("runtime")
    rs = np.random.RandomState(3)
    A = rs.normal(size=(30, 4, 5, 5))
    S = A @ np.swapaxes(A, -1, -2)
    S[3, 1] = np.diag([1, 1, 0, 1, 1.0])
    S[7, 2] = -np.eye(5)
    S[9, 0] = 1e15 * np.ones((5, 5))
    ok = smp._positive_definite(S)
    np.testing.assert_array_equal(ok, np.linalg.eigvalsh(S).min(axis=-1) > 0)
    R = smp._regularize(S.copy(), 1e-6)
    np.testing.assert_allclose(R, np.swapaxes(R, -1, -2), rtol=1e-12, atol=0)     # symmetric up to round-off
    w = np.linalg.eigvalsh(R)
    assert np.all(w[..., 0] >= np.maximum(1e-6, 1e-12 * w[..., -1]) * (1 - 1e-3))
    np.testing.assert_array_equal(R[ok], (0.5 * (S + np.swapaxes(S, -1, -2)))[ok])

    lib = rt.load_library()
    Q, C, dim = 2, 2, 3
    x = np.full((Q, C, dim), 0.5)
    cov = np.broadcast_to(np.eye(dim), (Q, C, dim, dim)).copy()
    cov[1, 1] = -np.eye(dim)                                            # not positive definite
    xi = np.zeros((Q, C, dim))
    xi[0, 0] = [-1.0, 0.0, 0.0]                                         # x_new = (-0.5, .5, .5): below lb
    xi[0, 1] = [0.25, 0.5, -0.25]
    lb, ub = np.zeros(dim), np.ones(dim)
    x_new, x_eval, inside = np.empty_like(x), np.empty_like(x), np.empty((Q, C), np.uint8)
    p = lambda a: a.ctypes.data
    bad = lib.pv_sampler_propose(Q, C, dim, p(cov), p(x), p(xi), p(lb), p(ub), p(x_new), p(inside), p(x_eval))
    assert bad == 1
    np.testing.assert_array_equal(inside, [[0, 1], [1, 1]])
    np.testing.assert_allclose(x_new[0, 0], [-0.5, 0.5, 0.5])
    np.testing.assert_allclose(x_eval[0, 0], x[0, 0])                   # outside: the objective is asked for x
    np.testing.assert_allclose(x_eval[0, 1], [0.75, 1.0, 0.25])
    np.testing.assert_allclose(x_new[1, 1], x[1, 1])                    # failed Cholesky: proposal = x
    assert lib.pv_sampler_propose(Q, C, 17, p(cov), p(x), p(xi), p(lb), p(ub), p(x_new), p(inside), p(x_eval)) < 0
    assert "dim" in rt.last_error()
    assert lib.pv_sampler_propose(Q, C, dim, None, p(x), p(xi), p(lb), p(ub), p(x_new), p(inside), p(x_eval)) < 0
    assert "pv_sampler_propose" in rt.last_error()


def test_library_rng_is_numpys_legacy_stream():
    """pv_rng_sampler_block / pv_sampler_advance_mt draw the sampler's blocks of random numbers in compiled code: bit for bit
    the numbers numpy's RandomState yields in the numpy loop's draw order (normal / uniform interleaved, cached second
    deviate of the polar method included), and the exported state continues the stream."""
    import ctypes
    rt = pkg("runtime")
    lib = rt.load_library()
    for n, Q, C, dim, seed in ((37, 5, 4, 3, 7), (10, 3, 1, 2, 1234), (60, 50, 4, 4, 99)):
        rs = np.random.RandomState(seed)
        rs.normal(size=3)                                   # odd count: a cached deviate is pending at hand-over
        st = rt.MT19937State.from_random_state(rs)
        S = max(C - 1, 1)
        xi, ua, us = np.empty((n, Q, C, dim)), np.empty((n, Q, C)), np.zeros((n, Q, S))
        for k in range(n):
            xi[k] = rs.normal(size=(Q, C, dim))
            ua[k] = rs.uniform(size=(Q, C))
            for j in range(C - 1, 0, -1):
                us[k, :, j - 1] = rs.uniform(size=Q)
        xi2, ua2, us2 = np.empty_like(xi), np.empty_like(ua), np.empty_like(us)
        assert lib.pv_rng_sampler_block(ctypes.addressof(st), n, Q, C, dim, xi2.ctypes.data, ua2.ctypes.data, us2.ctypes.data) == 0
        assert np.array_equal(xi, xi2) and np.array_equal(ua, ua2) and np.array_equal(us, us2)
        rs2 = np.random.RandomState(0)
        st.to_random_state(rs2)
        assert rs2.normal() == rs.normal() and rs2.uniform() == rs.uniform()
    assert lib.pv_rng_sampler_block(None, 1, 1, 1, 1, xi2.ctypes.data, ua2.ctypes.data, us2.ctypes.data) < 0


def test_batched_effective_sample_size_equals_the_per_chain_statement():
    smp = pkg("optimization.sampler")
    rs = np.random.RandomState(0)
    tr = np.zeros((6, 2, 2001, 2))
    for q in range(6):
        for i in range(1, 2001):
            tr[q, :, i] = (0.5 + 0.08 * q) * tr[q, :, i - 1] + rs.normal(size=(2, 2))
    tr[5, 0, :, 1] = 3.0                                       # a constant component: tau = n
    res = smp.SampleResult(tr, None, None, None, None, None, 0)
    ess = smp.effective_sample_size(res)
    for q in range(6):
        ch = tr[q, 0]
        b = smp.geweke_burn_in(ch)
        rest = ch[b:]
        tau = max(smp.autocorrelation_time_sokal(rest[:, p]) for p in range(2))
        assert res.burn_in[q] == b and ess[q] == pytest.approx(len(rest) / (1.0 + tau), rel=1e-12)


def test_results_table_matches_the_per_problem_frames():
    """GpuPetabSampler.results_table (all problems of a batch at once, one sort) against results_df problem by problem:
    same rows, same order, same columns but ``values``; order statistics (median, HDI bounds) identical, moments to
    round-off; also with the beta = 1 chain only and an even pooled sample count."""
    import types
    import pandas as pd
    drv = pkg("optimization.driver")
    rs = np.random.RandomState(5)
    for Q, C, n, pooled in ((7, 4, 301, True), (3, 3, 50, False), (2, 1, 4, True)):
        names = ["CL_MALE", "CL_FEMALE", "k_abs_MALE"]
        d = drv.GpuPetabSampler.__new__(drv.GpuPetabSampler)
        d.objective = types.SimpleNamespace(x_names=names, Q=Q, dim=len(names))
        d.uids = [f"xp{q:05d}" for q in range(Q)]
        d.optim_duration, d.sampler_duration = 1.5, 2.5
        d.sample_result = types.SimpleNamespace(trace_x=rs.lognormal(size=(Q, C, n, len(names))),
                                                effective_sample_size=rs.uniform(1, 100, size=Q))
        settings = {"n_samples": n - 1}
        tab = d.results_table(settings, pool_tempered_chains=pooled)
        ref = pd.concat([d.results_df(q, settings, pool_tempered_chains=pooled).drop(columns=["values"]) for q in range(Q)],
                        ignore_index=True)
        assert list(tab.columns) == list(ref.columns) and len(tab) == len(ref)
        for c in tab.columns:
            if tab[c].dtype.kind == "f":
                np.testing.assert_allclose(tab[c].to_numpy(), ref[c].to_numpy(), rtol=1e-13, atol=0, err_msg=c)
            else:
                assert (tab[c].to_numpy() == ref[c].to_numpy()).all(), c
        for c in ("median", "hdi_low", "hdi_high"):
            assert (tab[c].to_numpy() == ref[c].to_numpy()).all(), c
