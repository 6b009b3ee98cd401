"""GPU tests of the estimation stack built on the objective op: FIM, multi-problem batches, hierarchical form,
PEtab directories, multistart optimisation, adaptive-Metropolis/parallel-tempering sampling, results TSV."""
import json

import numpy as np
import pandas as pd
import pytest

from conftest import lognormal_draws, pkg

pytestmark = pytest.mark.gpu
OBS = ["y_cent", "y_gut", "y_peri"]


def _pk_groups(oracle, t, truths, n_ind, cv, seed):
    po = pkg("optimization.petab_optimization")
    cols = [oracle.SIMPLE_PK.states.index(o) for o in OBS]
    rs = np.random.RandomState(seed)
    groups = []
    for gid, (ka, cl) in truths.items():
        f = oracle.simple_pk_exact({"k_abs": ka, "CL": cl}, t)[:, cols]
        groups.append(po.GroupData(gid, f[None] * (1 + cv * rs.normal(size=(n_ind,) + f.shape))))
    return groups


def test_fim_and_problem_batch_vs_oracle(cuda_lib, rt, oracle):
    po = pkg("optimization.petab_optimization")
    t = np.linspace(0, 30, 12)
    cols = [oracle.SIMPLE_PK.states.index(o) for o in OBS]
    truths = {"MALE": (0.8, 1.1), "FEMALE": (1.3, 0.6)}
    problems = [_pk_groups(oracle, t, truths, n, cv, s) for n, cv, s in ((3, 0.05, 1), (10, 0.2, 2), (1, 0.0, 3))]
    priors = [{"k_abs_MALE": (0.5, 1.0), "CL_FEMALE": (0.5, 2.0)}, {}, {"CL_MALE": (1.0, 0.3)}]
    sigma = 0.1
    batch = po.PetabObjectiveBatch("simple_pk", ["k_abs", "CL"], problems, t, OBS, sigma=sigma, priors=priors, rtol=1e-9, atol=1e-13)
    X = np.stack([np.stack([lognormal_draws(10 + q, 4), lognormal_draws(20 + q, 4), lognormal_draws(30 + q, 4),
                            lognormal_draws(40 + q, 4)], axis=1) for q in range(3)])           # [Q, B, 4]
    o = batch.evaluate(X, order=2)
    for q in range(3):
        single = po.PetabObjective("simple_pk", ["k_abs", "CL"], problems[q], t, OBS, sigma=sigma, priors=priors[q], rtol=1e-9, atol=1e-13)
        lp, dlp = single.logp_and_grad(X[q])
        np.testing.assert_allclose(o["llh"][q] + o["lprior"][q], lp, rtol=1e-10)
        np.testing.assert_allclose(o["dllh"][q] + o["dlprior"][q], dlp, rtol=1e-8, atol=1e-8 * np.abs(dlp).max())
    # FIM against the oracle for one point: sum over groups of n_ind * s^T s / sigma^2, block diagonal per group
    q, b = 1, 2
    F = np.zeros((4, 4))
    for gi in range(2):
        over = {"k_abs": X[q, b, gi], "CL": X[q, b, 2 + gi]}
        s = oracle.simple_pk_exact_sens(over, t)[:, cols, :].reshape(-1, 2)
        Fg = 10 * s.T @ s / sigma**2
        for a, ia in ((0, gi), (1, 2 + gi)):
            for c, ic in ((0, gi), (1, 2 + gi)):
                F[ia, ic] = Fg[a, c]
    np.testing.assert_allclose(o["fim"][q, b], F, rtol=1e-6, atol=1e-8 * np.abs(F).max())


def test_hierarchical_objective_vs_oracle(cuda_lib, rt, oracle):
    po = pkg("optimization.petab_optimization")
    t = np.linspace(0, 40, 10)
    n_ind, C = 37, 3
    th_true = np.stack([lognormal_draws(1, n_ind), lognormal_draws(2, n_ind)], axis=1)
    rs = np.random.RandomState(5)
    y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in th_true])
    y = y * (1 + 0.05 * rs.normal(size=y.shape))
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-9, atol=1e-13)
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(t)
    pb.set_observables(["y_gut", "y_cent", "y_peri"])
    pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), 0.05)
    H = po.HierarchicalObjective(pb, n_ind, mu_prior=[(0.0, 1.0), (-0.3, 0.5)], omega_prior=[1.0, 0.7])
    theta = np.stack([np.stack([lognormal_draws(10 + c, n_ind), lognormal_draws(20 + c, n_ind)], axis=1) for c in range(C)])
    mu = rs.normal(size=(C, 2)) * 0.3
    omega = 0.5 + rs.rand(C, 2)
    lp, dth, dmu, dom = H(theta, mu, omega)
    for c in range(C):
        ref = 0.0
        for i in range(n_ind):
            ref += oracle.loglik_normal(oracle.simple_pk_exact({"k_abs": theta[c, i, 0], "CL": theta[c, i, 1]}, t), y[i], 0.05)
        pop, d_theta, d_mu, d_sigma = oracle.population_lognormal_logpdf(theta[c], mu[c], omega[c])
        ref += pop
        ref += oracle.log_prior_normal(mu[c], [0.0, -0.3], [1.0, 0.5])[0]
        ref += np.sum(0.5 * np.log(2 / np.pi) - np.log([1.0, 0.7]) - 0.5 * (omega[c] / np.array([1.0, 0.7])) ** 2)
        assert lp[c] == pytest.approx(ref, rel=1e-7)
        np.testing.assert_allclose(dmu[c], d_mu + oracle.log_prior_normal(mu[c], [0.0, -0.3], [1.0, 0.5])[1], rtol=1e-9)
        np.testing.assert_allclose(dom[c], d_sigma - omega[c] / np.array([1.0, 0.7]) ** 2, rtol=1e-9)
        i = 4
        s = oracle.simple_pk_exact_sens({"k_abs": theta[c, i, 0], "CL": theta[c, i, 1]}, t)
        g = oracle.loglik_normal_grad(oracle.simple_pk_exact({"k_abs": theta[c, i, 0], "CL": theta[c, i, 1]}, t), s, y[i], 0.05)
        np.testing.assert_allclose(dth[c, i], g + d_theta[i], rtol=1e-6)


@pytest.mark.parametrize("columns", [["k_abs", "CL"], ["CL", "k_abs"]])
def test_hierarchical_objective_device_resident(cuda_lib, rt, oracle, columns):
    """pv_hier_logp_grad + pv_hier_finish (CUDA tensors in and out, no host round trip) against the oracle and against the
    host-algebra ``HierarchicalObjective``; columns in and against the compiled sensitivity order; repeated evaluations are
    bit-identical (fixed-order reductions); theta <= 0 and a rank without individuals behave."""
    import torch
    po = pkg("optimization.petab_optimization")
    t = np.linspace(0, 40, 10)
    n_ind, C = 300, 5           # > 256: every thread of the epilogue CTA walks more than one individual
    th_true = np.stack([lognormal_draws(1, n_ind), lognormal_draws(2, n_ind)], axis=1)
    rs = np.random.RandomState(5)
    y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in th_true])
    y = y * (1 + 0.05 * rs.normal(size=y.shape))
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-9, atol=1e-13)
    pb.set_columns(columns)
    pb.set_timepoints(t)
    pb.set_observables(["y_gut", "y_cent", "y_peri"])
    pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), 0.05)
    mu_prior, omega_prior = [(0.0, 1.0), (-0.3, 0.5)], [1.0, 0.7]
    theta = np.stack([np.stack([lognormal_draws(10 + c, n_ind), lognormal_draws(20 + c, n_ind)], axis=1) for c in range(C)])
    mu = rs.normal(size=(C, 2)) * 0.3
    omega = 0.5 + rs.rand(C, 2)
    H = po.HierarchicalObjective(pb, n_ind, mu_prior=mu_prior, omega_prior=omega_prior)
    lp_h, dth_h, dmu_h, dom_h = H(theta, mu, omega)                      # host algebra (oracle-checked above)
    D = po.HierarchicalObjectiveDevice(pb, C, n_ind, mu_prior=mu_prior, omega_prior=omega_prior)
    th_d = torch.from_numpy(np.ascontiguousarray(theta.transpose(2, 0, 1))).cuda()
    mu_d, om_d = torch.from_numpy(mu).cuda(), torch.from_numpy(omega).cuda()
    red, dth = D(th_d, mu_d, om_d)
    torch.cuda.synchronize()
    red1, dth1 = red.cpu().numpy().copy(), dth.cpu().numpy().copy()
    assert int(D.status.abs().sum()) == 0
    np.testing.assert_allclose(red1[:, 0], lp_h, rtol=1e-12)
    np.testing.assert_allclose(red1[:, 1:3], dmu_h, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(red1[:, 3:5], dom_h, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dth1.transpose(1, 2, 0), dth_h, rtol=1e-12, atol=1e-12 * np.abs(dth_h).max())
    # ... and directly against the oracle for one chain
    c = 2
    names = {n: j for j, n in enumerate(columns)}
    ref = 0.0
    for i in range(n_ind):
        over = {n: theta[c, i, j] for n, j in names.items()}
        ref += oracle.loglik_normal(oracle.simple_pk_exact(over, t), y[i], 0.05)
    pop, d_theta, d_mu, d_sigma = oracle.population_lognormal_logpdf(theta[c], mu[c], omega[c])
    m0, s0 = np.array([a for a, _ in mu_prior]), np.array([b for _, b in mu_prior])
    ref += pop + oracle.log_prior_normal(mu[c], m0, s0)[0]
    ref += np.sum(0.5 * np.log(2 / np.pi) - np.log(omega_prior) - 0.5 * (omega[c] / np.array(omega_prior)) ** 2)
    assert red1[c, 0] == pytest.approx(ref, rel=1e-7)
    np.testing.assert_allclose(red1[c, 1:3], d_mu + oracle.log_prior_normal(mu[c], m0, s0)[1], rtol=1e-9)
    np.testing.assert_allclose(red1[c, 3:5], d_sigma - omega[c] / np.array(omega_prior) ** 2, rtol=1e-9)
    # deterministic
    red, dth = D(th_d, mu_d, om_d)
    torch.cuda.synchronize()
    assert np.array_equal(red.cpu().numpy(), red1) and np.array_equal(dth.cpu().numpy(), dth1)
    # theta <= 0 poisons exactly its chain
    bad = th_d.clone()
    bad[0, 1, 7] = -0.5
    red, _ = D(bad, mu_d, om_d)
    torch.cuda.synchronize()
    r = red.cpu().numpy()
    assert np.isnan(r[1, 0]) and np.all(np.isfinite(np.delete(r[:, 0], 1)))
    # a rank without individuals contributes zeros before the hyper-priors
    E = po.HierarchicalObjectiveDevice(pb, C, n_ind, mu_prior=mu_prior, omega_prior=omega_prior, first_individual=n_ind, n_local=0)
    pb.hier_logp_grad(C, 0, None, mu_d, om_d, None, E.red)
    torch.cuda.synchronize()
    assert np.all(E.red.cpu().numpy() == 0.0)


def test_petab_directory_literal_and_intended(cuda_lib, rt, oracle, tmp_path):
    io = pkg("optimization.petab_io")
    po = pkg("optimization.petab_optimization")
    P = pkg()
    t = np.linspace(0, 100, 6)
    groups = _pk_groups(oracle, t, {"MALE": (0.8, 1.1), "FEMALE": (1.3, 0.6)}, 4, 0.05, 9)
    priors = {"k_abs_MALE": (0.5, 1.0), "k_abs_FEMALE": (1.0, 1.0), "CL_MALE": (1.0, 1.0), "CL_FEMALE": (0.5, 1.0)}
    y = io.write_petab_like_reference(tmp_path / "uid1", P.MODELS["simple_pk"], groups, t, OBS, ["k_abs", "CL"], priors)
    lit = io.load_petab_problem(y, mode="literal")
    assert lit.x_names == ["k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE"] and np.all(lit.ub == 1e8)
    x = np.array([0.9, 1.2, 1.0, 0.7])
    # literal: all observed species start at 0 -> model output 0 -> likelihood constant in x (quirk Q4)
    const = sum(-oracle.loglik_normal(np.zeros_like(g.y[0]), g.y, 1.0) for g in groups)
    lp0 = oracle.log_prior_normal(x, [0.5, 1.0, 1.0, 0.5], [1.0] * 4)[0]
    assert lit.fun(x) == pytest.approx(const - lp0, rel=1e-9)
    it = io.load_petab_problem(y, mode="intended")
    direct = po.PetabObjective("simple_pk", ["k_abs", "CL"], groups, t, OBS, sigma=1.0, priors=priors)
    assert it.fun(x) == pytest.approx(direct.fun(x), rel=1e-12)


def test_multistart_optimizer_recovers_parameters(cuda_lib, rt, oracle):
    po = pkg("optimization.petab_optimization")
    drv = pkg("optimization.driver")
    t = np.linspace(0, 30, 16)
    truths = [{"MALE": (0.8, 1.1), "FEMALE": (1.3, 0.6)}, {"MALE": (2.0, 0.4), "FEMALE": (0.5, 1.8)}]
    problems = [_pk_groups(oracle, t, tr, 2, 0.0, 1) for tr in truths]
    batch = po.PetabObjectiveBatch("simple_pk", ["k_abs", "CL"], problems, t, OBS, sigma=0.01, rtol=1e-9, atol=1e-13)
    d = drv.GpuPetabSampler(batch, start_box=(np.full(4, 0.05), np.full(4, 5.0)))
    res = d.run_optimization(n_starts=24, maxiter=100, seed=7)
    for q, tr in enumerate(truths):
        want = np.array([tr["MALE"][0], tr["FEMALE"][0], tr["MALE"][1], tr["FEMALE"][1]])
        np.testing.assert_allclose(res.x[q, 0], want, rtol=1e-5)
    assert (res.converged[:, 0]).all()


def test_device_resident_optimizer_takes_the_host_loops_steps(cuda_lib, rt, oracle):
    """``optimizer.minimize_device`` (pv_optimizer_create / run / result: the multistart trust-region loop on the GPU) against
    ``optimizer.minimize`` (numpy loop, one objective launch per iteration) from the same start points: per start the same
    number of iterations and the same convergence flag, final iterates and objective values equal up to round-off; priors,
    bounds that bind, starts in stiff corners of the box [0, 1e8] (failed / re-run evaluations) included."""
    po = pkg("optimization.petab_optimization")
    om = pkg("optimization.optimizer")
    drv = pkg("optimization.driver")
    t = np.linspace(0, 30, 16)
    truths = [{"MALE": (0.8, 1.1), "FEMALE": (1.3, 0.6)}, {"MALE": (2.0, 0.4), "FEMALE": (0.5, 1.8)}, {"MALE": (1.0, 1.0), "FEMALE": (1.0, 1.0)}]
    problems = [_pk_groups(oracle, t, tr, 3, 0.05, 1 + k) for k, tr in enumerate(truths)]
    priors = [{"k_abs_MALE": (0.5, 1.0), "CL_FEMALE": (0.5, 2.0)}, {}, {"k_abs_MALE": (1.0, 0.1), "k_abs_FEMALE": (1.0, 0.1), "CL_MALE": (1.0, 0.1), "CL_FEMALE": (1.0, 0.1)}]
    batch = po.PetabObjectiveBatch("simple_pk", ["k_abs", "CL"], problems, t, OBS, sigma=0.05, priors=priors, rtol=1e-9, atol=1e-13)
    batch.lb = np.array([0.0, 0.0, 0.7, 0.0])                   # a lower bound above problem 1's optimum for CL_MALE: it binds
    S = 20
    x0 = om.uniform_startpoints(S, np.full(4, 0.05), np.full(4, 5.0), Q=3, seed=3)
    x0[:, -2:] = om.uniform_startpoints(2, np.zeros(4), np.full(4, 1e8), Q=3, seed=4)      # two starts per problem anywhere in [0, 1e8]
    d = drv.GpuPetabSampler(batch)
    host = om.minimize(d._eval_opt, x0, batch.lb, batch.ub, maxiter=60)
    n0 = rt.launch_count()
    dev = om.minimize_device(batch, x0, maxiter=60)
    assert rt.launch_count() - n0 >= 3 * int(dev.n_iter.max())
    # both results are sorted by objective value; undo that through the (unique) start points
    for q in range(3):
        hi = {tuple(np.round(a, 12)): k for k, a in enumerate(host.x0[q])}
        perm = [hi[tuple(np.round(a, 12))] for a in dev.x0[q]]
        np.testing.assert_array_equal(dev.n_iter[q], host.n_iter[q][perm])
        np.testing.assert_array_equal(dev.converged[q], host.converged[q][perm])
        np.testing.assert_allclose(dev.fval[q], host.fval[q][perm], rtol=1e-9)
        np.testing.assert_allclose(dev.x[q], host.x[q][perm], rtol=1e-7, atol=1e-9)
    assert dev.n_evals == host.n_evals
    assert dev.converged[[0, 2], 0].all() and dev.x[1, 0, 2] == pytest.approx(0.7)       # the binding bound
    want = np.array([truths[0]["MALE"][0], truths[0]["FEMALE"][0], truths[0]["MALE"][1], truths[0]["FEMALE"][1]])
    np.testing.assert_allclose(dev.x[0, 0], want, rtol=0.2)


def test_sampler_posterior_matches_laplace(cuda_lib, rt, oracle):
    """simple_chain, one well-identified parameter per group: the sampled posterior must agree with the Laplace
    approximation (mean = optimum, sd = FIM^-1/2) that the optimiser's own quantities give."""
    po = pkg("optimization.petab_optimization")
    drv = pkg("optimization.driver")
    t = np.linspace(0, 10, 21)
    rs = np.random.RandomState(2)
    groups = []
    for gid, k1 in (("MALE", 0.6), ("FEMALE", 1.4)):
        f = oracle.simple_chain_exact(k1, t)
        groups.append(po.GroupData(gid, f[None] * (1 + 0.05 * rs.normal(size=(5,) + f.shape))))
    # priors as in the reference's exact_prior experiments (factories/simple_chain.py): Normal(loc, scale) on the lin
    # scale; they are what keeps the tempered chains (likelihood ^ beta, prior untempered) inside a proper region
    priors = [{"k1_MALE": (0.6, 1.0), "k1_FEMALE": (1.4, 1.0)}]
    batch = po.PetabObjectiveBatch("simple_chain", ["k1"], [groups], t, ["S1", "S2"], sigma=0.05, priors=priors, rtol=1e-9, atol=1e-13)
    d = drv.GpuPetabSampler(batch, start_box=(np.full(2, 0.1), np.full(2, 3.0)))
    res = d.run_optimization(n_starts=8, maxiter=100)
    xopt = res.x[0, 0]
    np.testing.assert_allclose(xopt, [0.6, 1.4], rtol=0.05)
    o2 = batch.evaluate(xopt[None, None, :], order=2)
    F = o2["fim"][0, 0] + np.diag(o2["prior_precision"][0, 0])
    sd_laplace = np.sqrt(np.diag(np.linalg.inv(F)))
    sr = d.bayesian_sampler(n_samples=4000, n_chains=4, seed=11)
    cold = sr.trace_x[0, 0, 500:]
    np.testing.assert_allclose(cold.mean(axis=0), xopt, atol=3 * sd_laplace.max() / np.sqrt(50))
    np.testing.assert_allclose(cold.std(axis=0), sd_laplace, rtol=0.3)
    assert sr.effective_sample_size[0] > 50


@pytest.mark.parametrize("use_graph", [True, False])
def test_device_resident_sampler_walks_the_host_loops_chains(cuda_lib, rt, oracle, use_graph):
    """``sampler.sample_device`` (pv_sampler_create / advance / result: the whole adaptive-Metropolis + parallel-tempering loop
    on the GPU, one recorded CUDA graph per iteration) against the host loop (``sampler.sample``: one objective launch per
    iteration, bookkeeping on the host) on the same PetabObjectiveBatch and the same random stream: identical accept /
    swap decisions, chains equal up to round-off.  Wide bounds [0, 1e8]: the hot chains visit stiff parameter values
    (RODAS4 re-run inside the recorded iteration) and leave the box (rejected proposals); several blocks of random numbers,
    the last one short."""
    po = pkg("optimization.petab_optimization")
    smp = pkg("optimization.sampler")
    t = np.linspace(0, 30, 12)
    truths = {"MALE": (0.8, 1.1), "FEMALE": (1.3, 0.6)}
    problems = [_pk_groups(oracle, t, truths, n, cv, s) for n, cv, s in ((3, 0.05, 1), (10, 0.2, 2), (1, 0.0, 3), (5, 0.1, 4), (2, 0.5, 5))]
    priors = [{"k_abs_MALE": (0.5, 1.0), "CL_FEMALE": (0.5, 2.0)}, {}, {"CL_MALE": (1.0, 0.3)},
              {"k_abs_MALE": (0.5, 1.0), "k_abs_FEMALE": (1.0, 1.0), "CL_MALE": (1.0, 1.0), "CL_FEMALE": (0.5, 1.0)}, {}]
    batch = po.PetabObjectiveBatch("simple_pk", ["k_abs", "CL"], problems, t, OBS, sigma=0.1, priors=priors)
    x0 = np.tile(np.array([0.8, 1.3, 1.1, 0.6]), (len(problems), 1)) * (1 + 0.05 * np.arange(len(problems)))[:, None]
    n = 230

    def evaluate(X):
        o = batch.evaluate(X, order=0)
        return o["llh"], o["lprior"]

    for chains in (4, 1):
        host = smp.sample(evaluate, x0, batch.lb, batch.ub, n_samples=n, n_chains=chains, seed=7)
        n0 = rt.launch_count()
        # library_rng: the blocks of random numbers drawn by the library (numpy's legacy stream restated) / by numpy here
        dev = smp.sample_device(batch, x0, n_samples=n, n_chains=chains, seed=7, rng_block=100, use_graph=use_graph,
                                library_rng=use_graph)
        assert rt.launch_count() - n0 >= 4 * n          # the iterations did run through this library's kernels
        np.testing.assert_array_equal(dev.acceptance_rate, host.acceptance_rate)
        np.testing.assert_array_equal(dev.swap_rate, host.swap_rate)
        # round-off: the device code contracts a*b+c into fma and has its own exp / log / pow, and the hot chains'
        # proposal covariances (variances up to 1e15, eigenvalue floor 1e-12 relative) amplify that in the Cholesky
        # factor - their positions agree to 1e-4 relative, the beta = 1 chain (the posterior sample) to 1e-8
        np.testing.assert_allclose(dev.trace_x[:, 0], host.trace_x[:, 0], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(dev.trace_neglogpost[:, 0], host.trace_neglogpost[:, 0], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(dev.trace_x, host.trace_x, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(dev.trace_neglogprior, host.trace_neglogprior, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(dev.betas, host.betas, rtol=1e-10)
        assert dev.trace_x.shape == (len(problems), chains, n + 1, 4)
        assert 0.02 < dev.acceptance_rate[:, 0].mean() < 0.95
        if chains == 4:
            assert dev.trace_x[:, -1].max() > 100.0       # the hottest chain did roam far outside the posterior's bulk
    # the problem is usable again after the sampler is gone
    assert np.all(np.isfinite(batch.evaluate(x0[:, None, :], order=0)["llh"]))


def test_optimize_experiments_writes_reference_results_tsv(cuda_lib, rt, oracle, tmp_path):
    io = pkg("optimization.petab_io")
    drv = pkg("optimization.driver")
    P = pkg()
    t = np.linspace(0, 30, 8)
    yamls = []
    for k, cv in enumerate((0.01, 0.1)):
        groups = _pk_groups(oracle, t, {"MALE": (0.8, 1.1), "FEMALE": (1.3, 0.6)}, 3, cv, 20 + k)
        priors = {"k_abs_MALE": (0.5, 1.0), "k_abs_FEMALE": (1.0, 1.0), "CL_MALE": (1.0, 1.0), "CL_FEMALE": (0.5, 1.0)}
        yamls.append(io.write_petab_like_reference(tmp_path / "xps" / f"uid{k}", P.MODELS["simple_pk"], groups, t, OBS, ["k_abs", "CL"], priors))
    ok = drv.optimize_experiments(tmp_path / "results", yamls, mode="intended", start_box=(np.full(4, 0.05), np.full(4, 5.0)),
                                  settings={"n_starts": 8, "n_samples": 300, "n_chains": 4, "maxiter": 50})
    assert ok == [True, True]
    df = pd.read_csv(tmp_path / "results" / "uid0_results.tsv", sep="\t")
    # the reference's result columns (petab_optimization.py:228-269)
    assert list(df.columns) == ["id", "group", "parameter", "pid", "mean", "std", "median", "ess", "n_samples", "optim_duration",
                                "sampler_duration", "values", "hdi_low", "hdi_high"]
    assert list(df["pid"]) == ["k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE"] and list(df["group"]) == ["MALE", "FEMALE"] * 2
    vals = np.array(json.loads(df["values"][0]))
    assert vals.shape == (4, 301) and (df["hdi_low"] <= df["median"]).all() and (df["median"] <= df["hdi_high"]).all()
    # caching: a second call does nothing (petab_optimization.py:287-292)
    assert drv.optimize_experiments(tmp_path / "results", yamls) == [True, True]
    # errors are written to <uid>_errors.tsv, not raised (:340-348)
    bad = tmp_path / "xps" / "uidbad" / "petab.yaml"
    bad.parent.mkdir(parents=True)
    bad.write_text("format_version: 2\n")
    assert drv.optimize_experiment(bad, tmp_path / "results") is False
    assert (tmp_path / "results" / "uidbad_errors.tsv").exists()


# ---- the reference's own integration test (tests/experiments/test_factories.py): all three models, timepoints {5, 10} ------
FACTORY = {   # sampling / exact-prior values of src/parvar/experiments/factories/{simple_chain,simple_pk,icg_body_flat}.py
    "simple_pk": dict(observables=["y_cent", "y_gut", "y_peri"], dosage=None, n=20,
                      groups={"MALE": {"CL": (1.0, 1.0), "k_abs": (0.5, 1.0)}, "FEMALE": {"CL": (0.5, 1.0), "k_abs": (1.0, 1.0)}}),
    "icg_body_flat": dict(observables=["Cre_plasma_icg", "Cgi_plasma_icg"], dosage={"IVDOSE_icg": 10.0}, n=100,
                          groups={"MALE": {"BW": (75.0, 10.0), "LI__ICGIM_Vmax": (0.0369598840327503, 0.01)},
                                  "FEMALE": {"BW": (65.0, 10.0), "LI__ICGIM_Vmax": (0.02947, 0.01)}}),
}


@pytest.mark.parametrize("model", ["simple_pk", "icg_body_flat"])
@pytest.mark.parametrize("timepoints", [5, 10])
def test_create_petab_for_experiment_like_reference(cuda_lib, rt, oracle, tmp_path, model, timepoints):
    """Sampling -> batched simulation -> noise -> PEtab files, against the oracle's restatement of the same flow
    (same global RNG stream: np.random.seed(1234), lognormal per (group, parameter), then per row randint/reseed/normal)."""
    wf = pkg("experiments.workflow")
    io = pkg("optimization.petab_io")
    spec = FACTORY[model]
    cv = 0.1                                                   # model_experiment_factory default noise_cvs (:643-645)
    groups = [wf.GroupSpec(id=g, sampling=pars, estimation=pars, n_samples=spec["n"], steps=timepoints - 1, cv=cv)
              for g, pars in spec["groups"].items()]
    xp = wf.ExperimentSpec(id="xp1", model=model, groups=groups, observables=spec["observables"], dosage=spec["dosage"])
    yaml_file, dsets, frames = wf.create_petab_for_experiment(xp, tmp_path)
    assert (tmp_path / "xp1" / "petab.yaml").exists()
    # ---- oracle flow -------------------------------------------------------------------------------------------------
    om = oracle.MODELS[model]
    samples = oracle.create_samples_parameters({g: {p: (spec["n"], sc, lo) for p, (lo, sc) in pars.items()}
                                                for g, pars in spec["groups"].items()})
    t = np.linspace(0, 100, timepoints)
    idx = [om.states.index(o) for o in spec["observables"]]
    for g, pars in spec["groups"].items():
        for p in pars:
            np.testing.assert_array_equal(frames[g][p].to_numpy(), samples[g][p])          # sampling stream: bit identical
        for i in range(spec["n"]):
            over = dict(spec["dosage"] or {})
            over.update({p: samples[g][p][i] for p in pars})
            clean = (oracle.simple_pk_exact(over, t) if model == "simple_pk" else oracle.simulate(model, over, t))[:, idx]
            noisy = oracle.add_noise(clean, cv)                                            # consumes the global stream like :201-213
            got = np.stack([dsets[g][f"[{o}]"].values[i] for o in spec["observables"]], axis=1)
            assert np.max(np.abs(got - noisy) / (2e-9 + 2e-6 * np.abs(noisy))) <= 1.0
    # ---- the written problem loads back into the objective; literal mode reproduces quirk Q4 ----------------------------
    obj = io.load_petab_problem(yaml_file, mode="intended", dosage=spec["dosage"])
    x = np.array([np.mean(samples[g][p]) for p in list(spec["groups"]["MALE"]) for g in spec["groups"]])
    assert np.isfinite(obj.fun(x))
    lit = io.load_petab_problem(yaml_file, mode="literal")
    assert lit.fun(x) != obj.fun(x)


def test_hmc_on_hierarchical_posterior(cuda_lib, rt, oracle):
    """BASELINE config 3 in its intended use: a gradient-based sampler on the hierarchical simple_pk posterior
    (per-individual k_abs_i, CL_i + population mu, omega), every leapfrog step = one launch over chains x individuals with
    forward sensitivities.  All positive quantities are sampled on the log scale (Jacobian terms included)."""
    po = pkg("optimization.petab_optimization")
    hm = pkg("optimization.hmc")
    t = np.linspace(0.5, 24, 10)
    n_ind, C, P = 24, 4, 2
    rs = np.random.RandomState(0)
    mu_true, om_true = np.array([0.0, -0.4]), np.array([0.3, 0.25])
    th_true = np.exp(mu_true + om_true * rs.normal(size=(n_ind, P)))
    y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in th_true])
    y = y * (1 + 0.05 * rs.normal(size=y.shape))
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-8, atol=1e-12)
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(t, t0=0.0)
    pb.set_observables(["y_gut", "y_cent", "y_peri"])
    pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), 0.02)
    H = po.HierarchicalObjective(pb, n_ind, mu_prior=[(0.0, 1.0), (0.0, 1.0)], omega_prior=[0.5, 0.5])
    dim = n_ind * P + 2 * P

    def unpack(X):
        lth = X[:, :n_ind * P].reshape(-1, n_ind, P)
        return lth, X[:, n_ind * P:n_ind * P + P], X[:, n_ind * P + P:]

    def lg(X):
        lth, mu, lom = unpack(X)
        theta, omega = np.exp(lth), np.exp(lom)
        lp, dth, dmu, dom = H(theta, mu, omega)
        # change of variables theta = exp(lth), omega = exp(lom): + sum lth + sum lom, chain rule on the gradients
        lp = lp + lth.sum(axis=(1, 2)) + lom.sum(axis=1)
        g = np.concatenate([(dth * theta + 1.0).reshape(X.shape[0], -1), dmu, dom * omega + 1.0], axis=1)
        return lp, g

    x0 = np.concatenate([np.tile(np.log(th_true).ravel(), (C, 1)) + 0.05 * rs.normal(size=(C, n_ind * P)),
                         np.zeros((C, P)), np.log(0.3) * np.ones((C, P))], axis=1)
    # gradient of the transformed density against central differences (one coordinate of each kind)
    l0, g0 = lg(x0[:1])
    for i in (0, n_ind * P, n_ind * P + P + 1):
        e = np.zeros(dim)
        e[i] = 1e-5
        fd = (lg(x0[:1] + e)[0][0] - lg(x0[:1] - e)[0][0]) / 2e-5
        assert g0[0, i] == pytest.approx(fd, rel=2e-5, abs=1e-6 * abs(l0[0]))
    res = hm.hmc(lg, x0, n_samples=150, n_warmup=150, n_leapfrog=10, step_size=0.02, seed=5)
    assert np.all(res.accept_rate > 0.5) and res.divergences.sum() <= 3
    _, mu_s, lom_s = unpack(res.samples.reshape(-1, dim))
    assert np.all(np.abs(mu_s.mean(axis=0) - np.log(th_true).mean(axis=0)) < 0.25)
    assert np.all(np.exp(lom_s).mean(axis=0) < 0.8)
    assert res.n_grad_evals >= 300 * 8
    # the no-U-turn sampler on the same posterior (north_star: NUTS calls logp + gradient at every step)
    nt = pkg("optimization.nuts")
    rn = nt.nuts(lg, x0, n_samples=120, n_warmup=150, step_size=0.02, max_depth=6, seed=6)
    assert np.all(rn.accept_rate > 0.55) and rn.divergences.sum() <= 6 and rn.tree_depth.max() <= 6
    _, mu_n, lom_n = unpack(rn.samples.reshape(-1, dim))
    assert np.all(np.abs(mu_n.mean(axis=0) - np.log(th_true).mean(axis=0)) < 0.25)
    assert np.all(np.abs(mu_n.mean(axis=0) - mu_s.mean(axis=0)) < 0.2)           # the two samplers agree


def test_device_resident_hmc_walks_the_numpy_chain(cuda_lib, rt, oracle):
    """optimization/hmc.py::hmc_device (pv_hmc_begin / _drift / _kick / _accept around the device-resident hierarchical
    objective) against the numpy loop ``hmc`` around the host objective: same seed -> the same chains up to round-off over a
    short run (momenta, jittered trajectory lengths, acceptance, dual averaging and the mass-matrix window included);
    then a longer run of the device loop alone recovers the population means like the host samplers do."""
    torch = pytest.importorskip("torch")
    po = pkg("optimization.petab_optimization")
    hm = pkg("optimization.hmc")
    t = np.linspace(0.5, 24, 10)
    n_ind, C, P = 24, 4, 2
    rs = np.random.RandomState(0)
    mu_true, om_true = np.array([0.0, -0.4]), np.array([0.3, 0.25])
    th_true = np.exp(mu_true + om_true * rs.normal(size=(n_ind, P)))
    y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in th_true])
    y = y * (1 + 0.05 * rs.normal(size=y.shape))
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-8, atol=1e-12)
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(t, t0=0.0)
    pb.set_observables(["y_gut", "y_cent", "y_peri"])
    pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), 0.02)
    priors = dict(mu_prior=[(0.0, 1.0), (0.0, 1.0)], omega_prior=[0.5, 0.5])
    H = po.HierarchicalObjective(pb, n_ind, **priors)
    Hd = po.HierarchicalObjectiveDevice(pb, C, n_ind, **priors)
    dim = n_ind * P + 2 * P

    def lg(X):
        lth = X[:, :n_ind * P].reshape(-1, n_ind, P)
        mu, lom = X[:, n_ind * P:n_ind * P + P], X[:, n_ind * P + P:]
        theta, omega = np.exp(lth), np.exp(lom)
        lp, dth, dmu, dom = H(theta, mu, omega)
        return lp + lth.sum(axis=(1, 2)) + lom.sum(axis=1), \
            np.concatenate([(dth * theta + 1.0).reshape(X.shape[0], -1), dmu, dom * omega + 1.0], axis=1)

    x0 = np.concatenate([np.tile(np.log(th_true).ravel(), (C, 1)) + 0.05 * rs.normal(size=(C, n_ind * P)),
                         np.zeros((C, P)), np.log(0.3) * np.ones((C, P))], axis=1)
    kw = dict(n_samples=6, n_warmup=24, n_leapfrog=5, step_size=0.02, seed=5)       # the mass window opens at iteration 12
    ref = hm.hmc(lg, x0, **kw)
    got = hm.hmc_device(Hd, x0, **kw)
    assert got.n_grad_evals == ref.n_grad_evals
    # (host and device objective sum the individuals in different orders, and 30 leapfrog trajectories amplify that round-off:
    # the mass matrix is a variance estimated from 12 states; measured agreement 1.6e-6 there, better elsewhere)
    np.testing.assert_allclose(got.step_size, ref.step_size, rtol=1e-5)
    np.testing.assert_allclose(got.inv_mass, ref.inv_mass, rtol=1e-4)
    np.testing.assert_allclose(got.accept_rate, ref.accept_rate, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(got.samples, ref.samples, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got.logp, ref.logp, rtol=1e-7, atol=1e-5)
    assert np.array_equal(got.divergences, ref.divergences)
    n0 = rt.launch_count()
    res = hm.hmc_device(Hd, x0, n_samples=150, n_warmup=150, n_leapfrog=10, step_size=0.02, seed=5)
    assert rt.launch_count() - n0 >= res.n_grad_evals * 5                            # drift, sensitivity, epilogue, prior, kick per step
    assert np.all(res.accept_rate > 0.5) and res.divergences.sum() <= 3
    mu_s, lom_s = res.samples[:, :, n_ind * P:n_ind * P + P].reshape(-1, P), res.samples[:, :, n_ind * P + P:].reshape(-1, P)
    assert np.all(np.abs(mu_s.mean(axis=0) - np.log(th_true).mean(axis=0)) < 0.25)
    assert np.all(np.exp(lom_s).mean(axis=0) < 0.8)
    assert np.all(np.isfinite(res.samples))
    pb.close()


def test_device_resident_nuts_walks_the_numpy_chain(cuda_lib, rt, oracle):
    """optimization/nuts.py::nuts_device (pv_nuts_begin / _start / _drift / _leaf / _merge / _finish around the device-resident
    hierarchical objective) against the numpy loop ``nuts`` around the host objective: same seed -> the same trees (depths,
    leapfrog counts), step sizes, mass matrix and samples up to round-off over a short run; then a longer run of the device loop
    alone recovers the population means like the host samplers do."""
    pytest.importorskip("torch")
    po = pkg("optimization.petab_optimization")
    nt = pkg("optimization.nuts")
    t = np.linspace(0.5, 24, 10)
    n_ind, C, P = 24, 4, 2
    rs = np.random.RandomState(0)
    mu_true, om_true = np.array([0.0, -0.4]), np.array([0.3, 0.25])
    th_true = np.exp(mu_true + om_true * rs.normal(size=(n_ind, P)))
    y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in th_true])
    y = y * (1 + 0.05 * rs.normal(size=y.shape))
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-8, atol=1e-12)
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(t, t0=0.0)
    pb.set_observables(["y_gut", "y_cent", "y_peri"])
    pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), 0.02)
    priors = dict(mu_prior=[(0.0, 1.0), (0.0, 1.0)], omega_prior=[0.5, 0.5])
    H = po.HierarchicalObjective(pb, n_ind, **priors)
    Hd = po.HierarchicalObjectiveDevice(pb, C, n_ind, **priors)
    dim = n_ind * P + 2 * P

    def lg(X):
        lth = X[:, :n_ind * P].reshape(-1, n_ind, P)
        mu, lom = X[:, n_ind * P:n_ind * P + P], X[:, n_ind * P + P:]
        theta, omega = np.exp(lth), np.exp(lom)
        lp, dth, dmu, dom = H(theta, mu, omega)
        return lp + lth.sum(axis=(1, 2)) + lom.sum(axis=1), \
            np.concatenate([(dth * theta + 1.0).reshape(X.shape[0], -1), dmu, dom * omega + 1.0], axis=1)

    x0 = np.concatenate([np.tile(np.log(th_true).ravel(), (C, 1)) + 0.05 * rs.normal(size=(C, n_ind * P)),
                         np.zeros((C, P)), np.log(0.3) * np.ones((C, P))], axis=1)
    kw = dict(n_samples=5, n_warmup=26, step_size=0.02, max_depth=5, seed=6)       # the mass window opens at iteration 13
    ref = nt.nuts(lg, x0, **kw)
    got = nt.nuts_device(Hd, x0, **kw)
    assert got.n_grad_evals == ref.n_grad_evals and np.array_equal(got.tree_depth, ref.tree_depth)
    assert np.array_equal(got.divergences, ref.divergences)
    np.testing.assert_allclose(got.step_size, ref.step_size, rtol=1e-5)
    np.testing.assert_allclose(got.inv_mass, ref.inv_mass, rtol=1e-4)
    np.testing.assert_allclose(got.accept_rate, ref.accept_rate, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(got.samples, ref.samples, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got.logp, ref.logp, rtol=1e-7, atol=1e-5)
    res = nt.nuts_device(Hd, x0, n_samples=120, n_warmup=150, step_size=0.02, max_depth=6, seed=6)
    assert np.all(res.accept_rate > 0.55) and res.divergences.sum() <= 6 and res.tree_depth.max() <= 6
    mu_n = res.samples[:, :, n_ind * P:n_ind * P + P].reshape(-1, P)
    assert np.all(np.abs(mu_n.mean(axis=0) - np.log(th_true).mean(axis=0)) < 0.25)
    assert np.all(np.isfinite(res.samples))
    pb.close()
