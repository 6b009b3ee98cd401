"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): trajectories rtol 1e-6 / atol 1e-9; logp and gradients 1e-6 relative.
"""
import numpy as np
import pytest

from conftest import lognormal_draws, parity_violation, pkg

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _problem(rt, model, columns, tgrid, observables, dosage=None, rtol=1e-7, atol=1e-10, method=0):
    pb = rt.Problem(model)
    pb.set_solver(method=method, rtol=rtol, atol=atol)
    for k, v in (dosage or {}).items():
        pb.set_value(k, v)
    pb.set_columns(columns)
    pb.set_timepoints(tgrid)
    pb.set_observables(observables)
    return pb


def test_simple_chain_exact(cuda_lib, rt, oracle):
    t = np.linspace(0, 10, 21)
    k1 = lognormal_draws(1, 257)
    pb = _problem(rt, "simple_chain", ["k1"], t, ["S1", "S2"])
    Y, nf = pb.simulate_host(k1[None, :], len(k1))
    assert nf == 0
    ref = np.stack([oracle.simple_chain_exact(k, t) for k in k1])   # [B, T, 2]
    assert parity_violation(Y, ref) <= 1.0


def test_simple_pk_c1_single_set(cuda_lib, rt, oracle):
    """BASELINE config C1: simple_pk, defaults, linspace(0,100,21) and (0,100,50)."""
    for T in (21, 50):
        t = np.linspace(0, 100, T)
        pb = _problem(rt, "simple_pk", [], t, ["y_gut", "y_cent", "y_peri"])
        Y, nf = pb.simulate_host(None, 1)
        assert nf == 0
        ref = oracle.simple_pk_exact({}, t)
        assert parity_violation(Y[0], ref) <= 1.0
    # known-answer values
    t = np.array([0.0, 5.0, 10.0])
    pb = _problem(rt, "simple_pk", [], t, ["y_gut", "y_cent", "y_peri"])
    Y, _ = pb.simulate_host(None, 1)
    np.testing.assert_allclose(Y[0, 1, :], [0.00673795, 0.06623389, 0.10043281], rtol=1e-6)
    np.testing.assert_allclose(Y[0, 2, :], [4.53999e-05, 9.80974e-03, 1.58271e-02], rtol=1e-5)


def test_simple_pk_batched_draws(cuda_lib, rt, oracle):
    """C2 at a size the oracle finishes in seconds: 2000 log-normal draws x 50 timepoints."""
    B, T = 2000, 50
    t = np.linspace(0, 100, T)
    np.random.seed(1234)
    k_abs = np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B)
    CL = np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B)
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["[y_gut]", "[y_cent]", "[y_peri]"])
    status = np.empty(B, np.int32)
    steps = np.empty((2, B), np.int32)
    Y, nf = pb.simulate_host(np.stack([k_abs, CL]), B, status=status, steps=steps)
    assert nf == 0 and not status.any()
    assert steps[0].min() > 10
    ref = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in zip(k_abs, CL)])
    assert parity_violation(Y, ref) <= 1.0


def test_device_pointer_call_matches_host_call(cuda_lib, rt):
    torch = _torch()
    B, T = 4096 + 17, 20
    t = np.linspace(0, 100, T)
    th = np.stack([lognormal_draws(5, B), lognormal_draws(6, B)])
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["y_cent", "y_gut"])
    Yh, nf = pb.simulate_host(th, B)
    P = torch.from_numpy(th).cuda()
    Y = torch.empty((B, T, 2), dtype=torch.float64, device="cuda")
    status = torch.empty(B, dtype=torch.int32, device="cuda")
    pb.simulate(P, Y, status=status, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert nf == 0 and int(status.abs().sum()) == 0
    assert np.array_equal(Y.cpu().numpy(), Yh)       # same kernel, same inputs: bit identical


def test_icg_rodas4_vs_oracle(cuda_lib, rt, oracle):
    """Stiff PBPK model, dose 10 mg, BW and Vmax sampled (C4 inputs), 16 individuals."""
    B = 16
    t = np.linspace(0, 100, 21)
    BW = lognormal_draws(11, B, 75.0, 10.0)
    Vm = lognormal_draws(12, B, 0.0369598840327503, 0.01)
    obs = ["Cre_plasma_icg", "Cgi_plasma_icg", "Cve_icg", "LI__icg", "Afeces_icg"]
    pb = _problem(rt, "icg_body_flat", ["BW", "LI__ICGIM_Vmax"], t, obs, dosage={"IVDOSE_icg": 10.0})
    Y, nf = pb.simulate_host(np.stack([BW, Vm]), B)
    assert nf == 0
    idx = [oracle.ICG_STATES.index(o) for o in obs]
    ref = np.stack([oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0, "BW": b, "LI__ICGIM_Vmax": v}, t)[:, idx]
                    for b, v in zip(BW, Vm)])
    assert parity_violation(Y, ref) <= 1.0
    # known-answer values at defaults
    pb2 = _problem(rt, "icg_body_flat", [], np.array([0.0, 5, 10, 15, 20]), ["Cre_plasma_icg"], dosage={"IVDOSE_icg": 10.0})
    Y2, _ = pb2.simulate_host(None, 1)
    np.testing.assert_allclose(Y2[0, 1:, 0], [1.25351163e-03, 3.13601386e-04, 7.80941642e-05, 1.94248499e-05], rtol=2e-6)


def _noisy_data(oracle_Y, cv, seed):
    rs = np.random.RandomState(seed)
    return oracle_Y * (1.0 + cv * rs.normal(size=oracle_Y.shape))


def test_simple_pk_logp_and_grad(cuda_lib, rt, oracle):
    """logp (normal + lognormal) and forward-sensitivity gradient vs oracle, 1e-6 relative."""
    T, B = 20, 64
    t = np.linspace(0, 15, T)
    opt = pkg("optimization.petab_optimization")
    k_abs, CL = lognormal_draws(21, B), lognormal_draws(22, B)
    for noise_model, sigma in ((rt.NOISE_NORMAL, 0.05), (rt.NOISE_LOGNORMAL, 0.3)):
        if noise_model == rt.NOISE_NORMAL:
            obs, tt, atol = ["y_gut", "y_cent", "y_peri"], t, 1e-13
        else:
            # the log-normal model needs a signal that stays far above atol: central/peripheral only, t > 0
            obs, tt, atol = ["y_cent", "y_peri"], t[1:], 1e-16
        cols = [oracle.SIMPLE_PK.states.index(o) for o in obs]
        truth = oracle.simple_pk_exact({}, tt)[:, cols]
        data = _noisy_data(truth, 0.05, 3)[None]
        pb = _problem(rt, "simple_pk", ["k_abs", "CL"], tt, obs, rtol=1e-9, atol=atol)
        pb.set_timepoints(tt, t0=0.0)     # measurements may start after t0 (PEtab: initial values hold at 0)
        pb.set_data(opt.sufficient_statistics(data, noise_model)[..., None], sigma, noise_model)
        lp, g, _, nf = pb.logp_host(np.stack([k_abs, CL]), grad=True)
        lp0, _, _, _ = pb.logp_host(np.stack([k_abs, CL]), grad=False)
        assert nf == 0
        ref_lp, ref_g = np.empty(B), np.empty((2, B))
        for b in range(B):
            over = {"k_abs": k_abs[b], "CL": CL[b]}
            f = oracle.simple_pk_exact(over, tt)[:, cols]            # exact solution from t = 0
            s = oracle.simple_pk_exact_sens(over, tt)[:, cols, :]
            if noise_model == rt.NOISE_NORMAL:
                ref_lp[b] = oracle.loglik_normal(f, data[0], sigma)
                ref_g[:, b] = oracle.loglik_normal_grad(f, s, data[0], sigma)
            else:
                ref_lp[b] = oracle.loglik_lognormal(f, data[0], sigma)
                ref_g[:, b] = oracle.loglik_lognormal_grad(f, s, data[0], sigma)
        np.testing.assert_allclose(lp, ref_lp, rtol=1e-6)
        np.testing.assert_allclose(lp0, ref_lp, rtol=1e-6)
        np.testing.assert_allclose(g, ref_g, rtol=1e-6, atol=1e-6 * np.abs(ref_g).max())


def test_simple_pk_sensitivity_trajectories(cuda_lib, rt, oracle):
    torch = _torch()
    T, B = 21, 33
    t = np.linspace(0, 100, T)
    th = np.stack([lognormal_draws(31, B), lognormal_draws(32, B)])
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["y_gut", "y_cent", "y_peri"], rtol=1e-9, atol=1e-12)
    P = torch.from_numpy(th).cuda()
    Y = torch.empty((B, T, 3), dtype=torch.float64, device="cuda")
    S = torch.empty((B, T, 3, 2), dtype=torch.float64, device="cuda")
    pb.simulate_sens(P, Y, S)
    torch.cuda.synchronize()
    for b in range(0, B, 8):
        over = {"k_abs": th[0, b], "CL": th[1, b]}
        ref = oracle.simple_pk_exact_sens(over, t)                 # [T, 3, 2]
        got = S[b].cpu().numpy()
        assert parity_violation(got, ref, rtol=1e-6, atol=1e-9) <= 1.0


def test_icg_logp_grad_vs_oracle(cuda_lib, rt, oracle):
    """Rosenbrock forward sensitivities on the PBPK model: gradient of a Gaussian likelihood."""
    B, T = 4, 11
    t = np.linspace(0, 50, T)
    obs = ["Cre_plasma_icg", "Cgi_plasma_icg"]
    idx = [oracle.ICG_STATES.index(o) for o in obs]
    sens = ["BW", "LI__ICGIM_Vmax"]
    truth = oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0}, t)[:, idx]
    data = _noisy_data(truth, 0.05, 7)[None]
    sigma = 1e-4
    opt = pkg("optimization.petab_optimization")
    BW = lognormal_draws(41, B, 75.0, 10.0)
    Vm = lognormal_draws(42, B, 0.0369598840327503, 0.01)
    pb = _problem(rt, "icg_body_flat", sens, t, obs, dosage={"IVDOSE_icg": 10.0}, rtol=1e-9, atol=1e-14)
    pb.set_data(opt.sufficient_statistics(data, rt.NOISE_NORMAL)[..., None], sigma, rt.NOISE_NORMAL)
    lp, g, _, nf = pb.logp_host(np.stack([BW, Vm]), grad=True)
    assert nf == 0
    for b in range(B):
        over = {"IVDOSE_icg": 10.0, "BW": BW[b], "LI__ICGIM_Vmax": Vm[b]}
        Yo, So = oracle.simulate_sens("icg_body_flat", over, t, sens)
        ref_lp = oracle.loglik_normal(Yo[:, idx], data[0], sigma)
        ref_g = oracle.loglik_normal_grad(Yo[:, idx], So[:, idx, :], data[0], sigma)
        np.testing.assert_allclose(lp[b], ref_lp, rtol=1e-6)
        np.testing.assert_allclose(g[:, b], ref_g, rtol=1e-6)


@pytest.mark.parametrize("noise", ["normal", "lognormal"])
@pytest.mark.parametrize("n_data", [1, 7])
def test_icg_sensitivity_layouts_agree(cuda_lib, rt, oracle, noise, n_data):
    """The two layouts of the Rosenbrock sensitivity kernels (one trajectory per lane / one warp per block of the
    augmented system, include/parvar_b200.h pv_problem_set_layout) integrate the same method: logp, gradient, Fisher
    information and step counts agree on a ragged batch, for pooled and per-individual data, both noise models."""
    B = 77
    nm = rt.NOISE_NORMAL if noise == "normal" else rt.NOISE_LOGNORMAL
    t = np.linspace(0, 40, 9)
    if nm == rt.NOISE_LOGNORMAL:
        t = t[1:]                      # log y needs y > 0: the observed compartments are empty at t = 0
    obs = ["Cgi_plasma_icg", "Cre_plasma_icg", "LI__icg"]
    idx = [oracle.ICG_STATES.index(o) for o in obs]
    tt = np.unique(np.concatenate([[0.0], t]))        # the oracle integrates from its first time point
    truth = oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0}, tt)[len(tt) - len(t):, idx]
    opt = pkg("optimization.petab_optimization")
    rs = np.random.RandomState(5)
    data = np.abs(truth[None] * (1 + 0.05 * rs.normal(size=(n_data,) + truth.shape))) + 1e-12
    stats = np.stack([opt.sufficient_statistics(data[i:i + 1], nm) for i in range(n_data)], axis=-1)
    sigma = [0.1, 0.2, 0.15] if nm == rt.NOISE_LOGNORMAL else [1e-4, 2e-4, 1e-3]
    P = np.stack([lognormal_draws(51, B, 75.0, 10.0), lognormal_draws(52, B, 0.0369598840327503, 0.01)])
    res = {}
    for layout in (rt.LAYOUT_LANE, rt.LAYOUT_SPLIT):
        pb = _problem(rt, "icg_body_flat", ["BW", "LI__ICGIM_Vmax"], t, obs, dosage={"IVDOSE_icg": 10.0}, rtol=1e-9, atol=1e-14)
        pb.set_layout(layout)
        pb.set_timepoints(t, t0=0.0)
        pb.set_data(stats, sigma, nm)
        lp, g, _, nf = pb.logp_host(P, grad=True)
        lp2, g2, f2, _, nf2 = pb.logp_fim_host(P)
        assert nf == 0 and nf2 == 0
        np.testing.assert_array_equal(lp, lp2)       # the FIM variant of a layout reproduces its logp / gradient exactly
        np.testing.assert_array_equal(g, g2)
        res[layout] = (lp, g, f2)
        pb.close()
    (lp_a, g_a, f_a), (lp_b, g_b, f_b) = res[rt.LAYOUT_LANE], res[rt.LAYOUT_SPLIT]
    assert np.all(np.isfinite(lp_b)) and np.all(np.isfinite(g_b)) and np.all(np.isfinite(f_b))
    np.testing.assert_allclose(lp_b, lp_a, rtol=1e-7)
    np.testing.assert_allclose(g_b, g_a, rtol=1e-6, atol=1e-7 * np.abs(g_a).max())
    np.testing.assert_allclose(f_b, f_a, rtol=1e-6, atol=1e-7 * np.abs(f_a).max())


def test_icg_layout_auto_rule_and_errors(cuda_lib, rt):
    pb = _problem(rt, "icg_body_flat", ["BW"], np.linspace(0, 10, 3), ["Cre_plasma_icg"], dosage={"IVDOSE_icg": 10.0})
    with pytest.raises(rt.ParvarB200Error):
        pb.set_layout(7)
    pb.set_layout(rt.LAYOUT_AUTO)
    pb.close()


def test_empty_and_ragged_batches(cuda_lib, rt):
    t = np.linspace(0, 10, 5)
    pb = _problem(rt, "simple_pk", ["k_abs"], t, ["y_cent"])
    Y, nf = pb.simulate_host(np.zeros((1, 0)), 0)
    assert Y.shape == (0, 5, 1) and nf == 0
    for B in (1, 31, 32, 33, 127, 129):
        Y, nf = pb.simulate_host(np.full((1, B), 0.7), B)
        assert nf == 0 and np.all(Y == Y[:1])     # identical inputs -> identical lanes
    # single time point = initial state only
    pb.set_timepoints([0.0])
    Y, nf = pb.simulate_host(np.full((1, 3), 0.7), 3)
    assert nf == 0 and np.all(Y == 0.0)


def test_failed_trajectories_are_reported(cuda_lib, rt):
    t = np.linspace(0, 100, 5)
    pb = _problem(rt, "simple_pk", ["k_abs"], t, ["y_cent"])
    pb.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-7, atol=1e-10, max_steps=10)
    status = np.empty(4, np.int32)
    Y, nf = pb.simulate_host(np.full((1, 4), 1.0), 4, status=status)
    assert nf == 4 and np.all(status == 1)
    pb.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-7, atol=1e-10)
    th = np.array([[1.0, np.nan, 1.0, 1.0]])
    Y, nf = pb.simulate_host(th, 4, status=status)
    assert nf == 1 and status[1] != 0 and status[0] == 0


@pytest.mark.parametrize("layout", ["lane", "split"])
def test_rosenbrock_failures_are_reported_per_trajectory(cuda_lib, rt, layout):
    """Stiff path, logp + gradient: a NaN parameter row or an exhausted step budget fails ITS trajectory (status != 0,
    logp = NaN, counted in the return value) and leaves its neighbours in the same CTA untouched - in the split layout
    32 trajectories share a CTA and its barriers, so a failed one must keep walking through them."""
    T = 6
    t = np.linspace(0, 30, T)
    obs = ["Cre_plasma_icg", "Cgi_plasma_icg"]
    B = 45
    BW = lognormal_draws(31, B, 75.0, 10.0)
    Vm = lognormal_draws(32, B, 0.0369598840327503, 0.01)
    stats = np.stack([np.ones((T, 2)), np.full((T, 2), 1e-3), np.full((T, 2), 1e-6)])[..., None]

    def run(P, max_steps=1_000_000):
        pb = _problem(rt, "icg_body_flat", ["BW", "LI__ICGIM_Vmax"], t, obs, dosage={"IVDOSE_icg": 10.0})
        pb.set_solver(method=rt.METHOD_RODAS4, rtol=1e-7, atol=1e-11, max_steps=max_steps)
        pb.set_layout(rt.LAYOUT_LANE if layout == "lane" else rt.LAYOUT_SPLIT)
        pb.set_data(stats, 1e-4)
        out = pb.logp_host(P, grad=True)
        pb.close()
        return out

    good = np.stack([BW, Vm])
    lp0, g0, _, nf0 = run(good)
    assert nf0 == 0 and np.all(np.isfinite(lp0)) and np.all(np.isfinite(g0))
    bad = good.copy()
    bad[0, 3] = np.nan
    bad[1, 40] = np.nan
    lp1, g1, _, nf1 = run(bad)
    assert nf1 == 2 and np.isnan(lp1[3]) and np.isnan(lp1[40])
    keep = np.ones(B, bool)
    keep[[3, 40]] = False
    np.testing.assert_array_equal(lp1[keep], lp0[keep])          # neighbours are bit-identical
    np.testing.assert_array_equal(g1[:, keep], g0[:, keep])
    lp2, g2, _, nf2 = run(good, max_steps=20)                    # nobody finishes in 20 steps
    assert nf2 == B and np.all(np.isnan(lp2))


def test_concurrent_problems_with_different_time_grids(cuda_lib, rt):
    """Host threads with their own Problem (the study driver runs one per time grid) launch the same kernels with
    different dynamic shared-memory sizes at the same time.  The opt-in shared-memory limit is a per-function attribute:
    it must not depend on who launched last (regression: 'kernel launch: invalid argument' from the Rosenbrock
    sensitivity kernel when a short-grid thread lowered the limit under a long-grid thread)."""
    from concurrent.futures import ThreadPoolExecutor
    obs = ["Cre_plasma_icg", "Cgi_plasma_icg"]
    B = 40
    P = np.stack([lognormal_draws(61, B, 75.0, 10.0), lognormal_draws(62, B, 0.0369598840327503, 0.01)])

    def make(T):
        t = np.linspace(0, 100, T)
        pb = _problem(rt, "icg_body_flat", ["BW", "LI__ICGIM_Vmax"], t, obs, dosage={"IVDOSE_icg": 10.0})
        stats = np.stack([np.ones((T, 2)), np.full((T, 2), 1e-3), np.full((T, 2), 1e-6)])[..., None]
        pb.set_data(stats, 1e-4)
        return pb

    grids = [2, 160, 5, 80, 3, 40]
    ref = {}
    for T in grids:
        pb = make(T)
        ref[T] = pb.logp_fim_host(P)
        pb.close()

    def work(T):
        pb = make(T)
        out = [pb.logp_fim_host(P) for _ in range(15)]
        pb.close()
        return T, out

    with ThreadPoolExecutor(max_workers=len(grids)) as pool:
        for T, outs in pool.map(work, grids):
            for lp, g, f, _, nf in outs:
                assert nf == 0
                np.testing.assert_array_equal(lp, ref[T][0])
                np.testing.assert_array_equal(g, ref[T][1])
                np.testing.assert_array_equal(f, ref[T][2])


def test_full_size_properties_c2(cuda_lib, rt):
    """BASELINE C2 at full size (1e6 draws x 50 timepoints): size-independent properties.
    simple_pk is linear: mass only leaves through clearance, so 0 <= y_gut+y_cent+y_peri <= 1 and y_gut is
    exp(-k_abs t) exactly; doubling the dose (y_gut(0)=2) doubles every trajectory (linearity)."""
    torch = _torch()
    B, T = 1_000_000, 50
    t = np.linspace(0, 100, T)
    np.random.seed(1234)
    k_abs = np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B)
    CL = np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B)
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["y_gut", "y_cent", "y_peri"])
    P = torch.from_numpy(np.stack([k_abs, CL])).cuda()
    Y = torch.empty((B, T, 3), dtype=torch.float64, device="cuda")
    status = torch.empty(B, dtype=torch.int32, device="cuda")
    pb.simulate(P, Y, status=status)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    total = Y.sum(dim=2)
    assert float(total.max()) <= 1.0 + 1e-9 and float(total.min()) >= -1e-9
    gut = torch.exp(-P[0][:, None] * torch.from_numpy(t).cuda()[None, :])
    assert float(((Y[:, :, 0] - gut).abs() / (1e-9 + 1e-6 * gut)).max()) <= 1.0
    pb.set_value("y_gut", 2.0)
    Y2 = torch.empty_like(Y)
    pb.simulate(P, Y2)
    torch.cuda.synchronize()
    assert float(((Y2 - 2 * Y).abs() / (2e-9 + 2e-6 * Y.abs())).max()) <= 1.0


def test_full_size_properties_icg(cuda_lib, rt, oracle):
    """BASELINE C4-sized forward batch of the stiff PBPK model (128 000 draws of BW and Vmax): size-independent
    properties.  The model conserves ICG: injected dose / Mr = sum of compartment amounts + remaining dose at every grid
    point (volumes depend on the sampled BW); concentrations stay non-negative; a trajectory does not depend on which
    lane integrated it (permuting the batch permutes the output bit for bit)."""
    torch = _torch()
    B, T = 128_000, 20
    t = np.linspace(0, 100, T)
    BW = lognormal_draws(21, B, 75.0, 10.0)
    Vm = lognormal_draws(22, B, 0.0369598840327503, 0.01)
    pb = _problem(rt, "icg_body_flat", ["BW", "LI__ICGIM_Vmax"], t, list(oracle.ICG_STATES), dosage={"IVDOSE_icg": 10.0})
    P = torch.from_numpy(np.stack([BW, Vm])).cuda()
    Y = torch.empty((B, T, len(oracle.ICG_STATES)), dtype=torch.float64, device="cuda")
    status = torch.empty(B, dtype=torch.int32, device="cuda")
    pb.simulate(P, Y, status=status)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    Yh = Y.cpu().numpy()
    par = dict(oracle.ICG_DEFAULTS)
    par["BW"] = BW
    v = oracle._icg_volumes_flows(par)
    vol = {"Afeces_icg": 1.0, "Car_icg": v["Var"], "Cgi_plasma_icg": v["Vgi_plasma"], "Chv_icg": v["Vhv"],
           "Cli_plasma_icg": v["Vli_plasma"], "Clu_plasma_icg": v["Vlu_plasma"], "Cpo_icg": v["Vpo"],
           "Cre_plasma_icg": v["Vre_plasma"], "Cve_icg": v["Vve"], "LI__icg": v["Vli_tissue"], "LI__icg_bi": v["Vbi"]}
    i = oracle.ICG_STATES.index
    amount = sum(Yh[:, :, i(sid)] * (np.asarray(V) * np.ones(B))[:, None] for sid, V in vol.items())
    amount = amount + Yh[:, :, i("IVDOSE_icg")] / par["Mr_icg"]
    np.testing.assert_allclose(amount, 10.0 / par["Mr_icg"], rtol=2e-6)
    assert Yh.min() >= -1e-9
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).cuda()
    Y2 = torch.empty_like(Y)
    pb.simulate(P[:, perm].contiguous(), Y2)
    torch.cuda.synchronize()
    assert torch.equal(Y2, Y[perm])


def test_ode_sample_simulator_drop_in(cuda_lib, oracle):
    """The reference-facing class: DataFrame in, (sim x time) dataset out, noise stream identical to the oracle's
    restatement of petab_factory.py:192-219."""
    import pandas as pd
    pf = pkg("experiments.petab_factory")
    ex = pkg("experiments.experiment")
    P = pkg()
    samples = pf.create_samples_parameters(
        {pf.Category.MALE: {pf.PKPDParameters.CL: pf.LognormParameters(n=20, sigma=1.0, mu=1.0),
                            pf.PKPDParameters.k_abs: pf.LognormParameters(n=20, sigma=1.0, mu=0.5)}})
    df = pd.DataFrame({par.value: v for par, v in samples[pf.Category.MALE].items()})
    np.testing.assert_allclose(df["CL"].values[:3], [1.04699267, 0.26233685, 2.33085045], rtol=1e-7)
    np.testing.assert_allclose(df["k_abs"].values[:3], [0.81277741, 0.65610198, 0.97999947], rtol=1e-7)
    sim = pf.ODESampleSimulator(P.MODELS["simple_pk"])
    obs = [ex.Observable(id="y_cent"), ex.Observable(id="y_gut"), ex.Observable(id="y_peri")]
    st = pf.SimulationSettings(start=0.0, end=100.0, steps=20, noise=ex.Noise(add_noise=False, cv=0.05), observables=obs)
    ds = sim.simulate_samples(df, st)
    assert list(ds.data_vars) == ["[y_cent]", "[y_gut]", "[y_peri]"] and obs[0].id == "[y_cent]"
    t = np.linspace(0, 100, 21)
    for i in (0, 7, 19):
        ref = oracle.simple_pk_exact({"CL": df["CL"][i], "k_abs": df["k_abs"][i]}, t)
        got = np.stack([ds["[y_gut]"].values[i], ds["[y_cent]"].values[i], ds["[y_peri]"].values[i]], axis=1)
        assert parity_violation(got, ref) <= 1.0
    # with noise: same global-RNG consumption as the reference loop
    st.noise = ex.Noise(add_noise=True, cv=0.05)
    np.random.seed(99)
    dsn = sim.simulate_samples(df, st)
    np.random.seed(99)
    clean = np.stack([ds[v].values for v in ds.data_vars], axis=2)     # [sim, time, obs]
    for i in range(20):
        noisy = oracle.add_noise(clean[i], 0.05)
        got = np.stack([dsn[v].values[i] for v in dsn.data_vars], axis=1)
        np.testing.assert_array_equal(got, noisy)


def test_c3_chains_times_individuals_logp_grad(cuda_lib, rt, oracle):
    """BASELINE config C3 shape: n_chains x n_individuals trajectories, every individual has its own data and its own
    (k_abs_i, CL_i); logp per chain = deterministic segment sum, gradient per individual.  Oracle-checked at
    8 x 200 (exact solution + Van Loan sensitivities), shape/consistency-checked at the full 8 x 1000."""
    torch = _torch()
    opt = pkg("optimization.petab_optimization")
    T = 20
    t = np.linspace(0, 100, T)
    obs = ["y_gut", "y_cent", "y_peri"]
    for n_chains, n_ind, check in ((8, 200, True), (8, 1000, False)):
        truth_th = np.stack([lognormal_draws(50, n_ind), lognormal_draws(51, n_ind)])
        rs = np.random.RandomState(52)
        y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in zip(*truth_th)])     # [n_ind, T, 3]
        y = y * (1.0 + 0.05 * rs.normal(size=y.shape))
        stats = np.stack([opt.sufficient_statistics(y[i:i + 1], rt.NOISE_NORMAL) for i in range(n_ind)], axis=-1)
        pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, obs, rtol=1e-9, atol=1e-13)
        pb.set_data(stats, 0.05, rt.NOISE_NORMAL)
        th = np.concatenate([np.stack([lognormal_draws(60 + c, n_ind), lognormal_draws(80 + c, n_ind)]) for c in range(n_chains)], axis=1)
        B = n_chains * n_ind
        P = torch.from_numpy(th).cuda()
        lp = torch.empty(B, dtype=torch.float64, device="cuda")
        g = torch.empty((2, B), dtype=torch.float64, device="cuda")
        seg = torch.empty(n_chains, dtype=torch.float64, device="cuda")
        status = torch.empty(B, dtype=torch.int32, device="cuda")
        pb.logp_grad(P, lp, g, seg_len=n_ind, logp_seg=seg, status=status)
        torch.cuda.synchronize()
        assert int(status.abs().sum()) == 0
        lp_h, g_h, seg_h = lp.cpu().numpy(), g.cpu().numpy(), seg.cpu().numpy()
        np.testing.assert_allclose(seg_h, lp_h.reshape(n_chains, n_ind).sum(axis=1), rtol=1e-12)
        # same call through the host-buffer entry point: bit identical
        lp2, g2, seg2, nf = pb.logp_host(th, seg_len=n_ind, grad=True)
        assert nf == 0 and np.array_equal(lp2, lp_h) and np.array_equal(g2, g_h) and np.array_equal(seg2, seg_h)
        if check:
            for c in (0, n_chains - 1):
                ref = np.empty(n_ind)
                for i in range(n_ind):
                    over = {"k_abs": th[0, c * n_ind + i], "CL": th[1, c * n_ind + i]}
                    f = oracle.simple_pk_exact(over, t)
                    ref[i] = oracle.loglik_normal(f, y[i], 0.05)
                    if i < 5:
                        s = oracle.simple_pk_exact_sens(over, t)
                        rg = oracle.loglik_normal_grad(f, s, y[i], 0.05)
                        np.testing.assert_allclose(g_h[:, c * n_ind + i], rg, rtol=1e-6, atol=1e-6 * np.abs(rg).max())
                np.testing.assert_allclose(lp_h[c * n_ind:(c + 1) * n_ind], ref, rtol=1e-6)
                np.testing.assert_allclose(seg_h[c], ref.sum(), rtol=1e-6)


def gradient_bar(got, ref, mass, rel=1e-6):
    """The 1e-6 gradient bar of the north star with the norm spelled out.  ``mass`` = sum over (t, k) of
    |(y - f)/sigma^2 * df/dtheta|, the absolute size of the terms a gradient component is the sum of.
      (A) |got - ref| <= rel * mass       for every component (what "1e-6 relative" can mean for a sum whose terms
                                           cancel: an integrator cannot deliver more digits than the terms carry);
      (B) |got - ref| <= rel * |ref|      component-wise for every component that cancels by less than 10x.
    Returns the two worst ratios (each must be <= 1)."""
    err = np.abs(got - ref)
    a = float(np.max(err / (rel * mass)))
    well = mass <= 10.0 * np.abs(ref)
    b = float(np.max((err / (rel * np.abs(ref)))[well])) if well.any() else 0.0
    return a, b


@pytest.mark.parametrize("which", ["C3", "C4"])
def test_bench_logp_grad_settings_meet_the_bar(cuda_lib, rt, oracle, which):
    """bench.py's ``logp_grad`` numbers are quoted at exactly these settings (``bench_cases``: model, T = 20, sigma,
    rtol / atol, seeded draws, 8 chains x 1000 individuals in one launch): logp within 1e-6 relative and the gradient
    within 1e-6 (``gradient_bar``) of the oracle at rtol 1e-12, on 64 (chain, individual) pairs; for the PBPK model in
    both kernel layouts."""
    bc = pkg("bench_cases")
    opt = pkg("optimization.petab_optimization")
    case = {"C3": bc.C3, "C4": bc.C4}[which]
    t = case.tgrid
    m = oracle.MODELS[case.model]
    idx = [m.states.index(o) for o in case.observables]
    truth = oracle.simulate(case.model, {**case.fixed, **case.mean_parameters()}, t)[:, idx]
    C, n_ind = bc.N_CHAINS, bc.N_INDIVIDUALS_PER_GPU
    y, theta = case.inputs(truth, n_ind)
    pb = _problem(rt, case.model, list(case.columns), t, list(case.observables), dosage=case.fixed,
                  rtol=case.rtol, atol=case.atol)
    pb.set_data(np.stack([opt.sufficient_statistics(y[i:i + 1], rt.NOISE_NORMAL) for i in range(n_ind)], axis=-1),
                case.sigma, rt.NOISE_NORMAL)
    P = np.ascontiguousarray(theta.transpose(2, 0, 1).reshape(len(case.columns), C * n_ind))
    rows = [pb.sens_ids.index(c) for c in case.columns]
    pairs = [(c, i) for c in range(C) for i in range(8)]
    ref = []
    for c, i in pairs:
        over = {**case.fixed, **{k: float(theta[c, i, p]) for p, k in enumerate(case.columns)}}
        if case.model == "simple_pk":
            f, s = oracle.simple_pk_exact(over, t)[:, idx], oracle.simple_pk_exact_sens(over, t)[:, idx, :]
        else:
            Yo, So = oracle.simulate_sens(case.model, over, t, list(case.columns))
            f, s = Yo[:, idx], So[:, idx, :]
        mass = np.abs(((y[i] - f) / case.sigma ** 2)[..., None] * s).sum(axis=(0, 1))
        ref.append((oracle.loglik_normal(f, y[i], case.sigma), oracle.loglik_normal_grad(f, s, y[i], case.sigma), mass))
    layouts = (rt.LAYOUT_AUTO, rt.LAYOUT_LANE) if case.model == "icg_body_flat" else (rt.LAYOUT_AUTO,)
    for layout in layouts:
        pb.set_layout(layout)
        lp, g, seg, nf = pb.logp_host(P, seg_len=n_ind, grad=True)
        assert nf == 0 and np.all(np.isfinite(lp)) and np.all(np.isfinite(g))
        worst_a = worst_b = 0.0
        for (c, i), (rl, rg, mass) in zip(pairs, ref):
            b = c * n_ind + i
            assert abs(lp[b] - rl) <= 1e-6 * abs(rl), (which, c, i, lp[b], rl)
            a_, b_ = gradient_bar(g[rows, b], rg, mass)
            worst_a, worst_b = max(worst_a, a_), max(worst_b, b_)
        assert worst_a <= 1.0 and worst_b <= 1.0, (which, layout, worst_a, worst_b)


def test_c5_sweep_problems_in_one_launch(cuda_lib, rt, oracle):
    """BASELINE config C5: many independent PEtab problems of one model (different sample counts, noise levels)
    evaluated in ONE launch - the problem index is just another data set.  Pooled statistics per (problem, group)."""
    opt = pkg("optimization.petab_optimization")
    T = 10
    t = np.linspace(0, 100, T)
    obs = ["y_gut", "y_cent", "y_peri"]
    rs = np.random.RandomState(7)
    problems = []
    for n_samples in (1, 2, 5, 20, 80):            # factories/__init__.py:12-20 "samples"
        for cv in (0.0, 0.05, 0.5):                 # "noise_cvs"
            th = np.stack([lognormal_draws(int(rs.randint(1 << 30)), n_samples), lognormal_draws(int(rs.randint(1 << 30)), n_samples)])
            y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in zip(*th)])
            problems.append(y * (1.0 + cv * rs.normal(size=y.shape)))
    n_prob = len(problems)
    stats = np.stack([opt.sufficient_statistics(y, rt.NOISE_NORMAL) for y in problems], axis=-1)          # [3,T,3,n_prob]
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, obs, rtol=1e-9, atol=1e-13)
    pb.set_data(stats, 1.0, rt.NOISE_NORMAL)       # sigma = 1 as the reference writes (quirk Q5)
    n_x = 7                                        # 7 candidate parameter vectors per problem
    X = np.stack([lognormal_draws(300, n_x), lognormal_draws(301, n_x)])                                   # [2, n_x]
    P = np.repeat(X, n_prob, axis=1)               # trajectory x*n_prob + q  -> data set q
    lp, g, _, nf = pb.logp_host(P, grad=True)
    assert nf == 0
    for x in (0, n_x - 1):
        over = {"k_abs": X[0, x], "CL": X[1, x]}
        f = oracle.simple_pk_exact(over, t)
        s = oracle.simple_pk_exact_sens(over, t)
        for q in (0, 4, n_prob - 1):
            np.testing.assert_allclose(lp[x * n_prob + q], oracle.loglik_normal(f, problems[q], 1.0), rtol=1e-6)
            rg = oracle.loglik_normal_grad(f, s, problems[q], 1.0)
            np.testing.assert_allclose(g[:, x * n_prob + q], rg, rtol=1e-6, atol=1e-6 * np.abs(rg).max())


def test_petab_objective_pypesto_conventions(cuda_lib, rt, oracle):
    """The objective mirror: negative log posterior, x ordered par-major/group-minor, priors, literal mode (quirk Q4)."""
    opt = pkg("optimization.petab_optimization")
    t = np.linspace(0, 100, 21)
    obs = ["y_cent", "y_gut", "y_peri"]                       # factories/simple_pk.py:4-17 order
    cols = [oracle.SIMPLE_PK.states.index(o) for o in obs]
    rs = np.random.RandomState(3)
    groups, ys = [], {}
    for gid, (ka, cl) in (("MALE", (0.8, 1.1)), ("FEMALE", (1.3, 0.6))):
        y = np.stack([oracle.simple_pk_exact({"k_abs": ka, "CL": cl}, t)[:, cols]] * 4)
        y = y * (1 + 0.05 * rs.normal(size=y.shape))
        groups.append(opt.GroupData(gid, y))
        ys[gid] = y
    priors = {"k_abs_MALE": (0.5, 1.0), "k_abs_FEMALE": (1.0, 1.0), "CL_MALE": (1.0, 1.0), "CL_FEMALE": (0.5, 1.0)}
    obj = opt.PetabObjective("simple_pk", ["k_abs", "CL"], groups, t, obs, sigma=1.0, priors=priors, rtol=1e-9, atol=1e-13)
    assert obj.x_names == ["k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE"]
    x = np.array([0.9, 1.2, 1.0, 0.7])
    fval, grad = obj(x, sensi_orders=(0, 1))
    ref, ref_g = 0.0, np.zeros(4)
    for gi, gid in enumerate(("MALE", "FEMALE")):
        over = {"k_abs": x[gi], "CL": x[2 + gi]}
        f = oracle.simple_pk_exact(over, t)[:, cols]
        s = oracle.simple_pk_exact_sens(over, t)[:, cols, :]
        ref -= oracle.loglik_normal(f, ys[gid], 1.0)
        gg = oracle.loglik_normal_grad(f, s, ys[gid], 1.0)
        ref_g[gi] -= gg[0]
        ref_g[2 + gi] -= gg[1]
    lp, glp = oracle.log_prior_normal(x, [0.5, 1.0, 1.0, 0.5], [1.0, 1.0, 1.0, 1.0])
    ref -= lp
    ref_g -= glp
    assert fval == pytest.approx(ref, rel=1e-6)
    np.testing.assert_allclose(grad, ref_g, rtol=1e-6, atol=1e-6 * np.abs(ref_g).max())
    assert obj.fun(x) == pytest.approx(fval, rel=1e-12)
    assert obj(x, (0,), return_dict=True)["fval"] == pytest.approx(fval, rel=1e-12)
    # batched entry = the single-point calls
    X = np.stack([x, x * 1.1, x * 0.7])
    L, G = obj.logp_and_grad(X)
    assert L[0] == pytest.approx(-fval, rel=1e-12) and np.allclose(G[0], -grad, rtol=1e-12)
    # literal mode (quirk Q4): observed species start at 0, no dose -> the model output is identically 0, the
    # likelihood is constant in x and the gradient comes from the prior alone
    lit = opt.PetabObjective("simple_pk", ["k_abs", "CL"], groups, t, obs, sigma=1.0, priors=priors, mode="literal")
    f1, g1 = lit(x, (0, 1))
    f2, g2 = lit(x * 2.0, (0, 1))
    const = sum(-oracle.loglik_normal(np.zeros_like(ys[g][0]), ys[g], 1.0) for g in ys)
    assert f1 == pytest.approx(const - oracle.log_prior_normal(x, [0.5, 1.0, 1.0, 0.5], [1.0] * 4)[0], rel=1e-9)
    np.testing.assert_allclose(g1, -oracle.log_prior_normal(x, [0.5, 1.0, 1.0, 0.5], [1.0] * 4)[1], rtol=1e-9, atol=1e-12)
    assert f2 != f1
    # failed integration -> inf / nan, not an exception
    obj.problem.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-9, atol=1e-13, max_steps=3)
    fbad, gbad = obj(x, (0, 1))
    assert np.isinf(fbad) and np.all(np.isnan(gbad))


def test_stiff_parameter_values_switch_to_rodas4(cuda_lib, rt, oracle):
    """METHOD_AUTO on the (non-stiff) simple_pk model: rows with huge rate constants - what a tempered chain or a
    uniform start point in the PEtab bounds [0, 1e8] produces - exhaust the explicit step budget and are re-run with
    RODAS4 in the same call; the other rows keep their DOPRI5 result bit for bit."""
    t = np.linspace(0, 100, 21)
    k_abs = np.array([1.0, 3e4, 0.5, 1e8, 2.0, 7e5])
    CL = np.array([1.0, 0.5, 2e6, 1.0, 0.3, 5e7])
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["y_gut", "y_cent", "y_peri"])
    status = np.empty(6, np.int32)
    steps = np.empty((2, 6), np.int32)
    Y, nf = pb.simulate_host(np.stack([k_abs, CL]), 6, status=status, steps=steps)
    assert nf == 0 and not status.any()
    assert steps.sum(axis=0).max() < 4000                      # nobody ground through millions of explicit steps
    ref = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in zip(k_abs, CL)])
    assert parity_violation(Y, ref) <= 1.0
    pb.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-7, atol=1e-10)
    Y2, _ = pb.simulate_host(np.stack([k_abs[[0, 2 - 2, 4]], CL[[0, 0, 4]]]), 3)
    assert np.array_equal(Y2[0], Y[0]) and np.array_equal(Y2[2], Y[4])
    # gradient mode goes through the same switch
    pb.set_solver(method=rt.METHOD_AUTO, rtol=1e-8, atol=1e-12)
    truth = oracle.simple_pk_exact({}, t)
    pb.set_data(np.stack([np.ones_like(truth), truth, truth * truth])[..., None], 0.1)
    lp, g, _, nf = pb.logp_host(np.stack([k_abs, CL]), grad=True)
    assert nf == 0 and np.all(np.isfinite(lp)) and np.all(np.isfinite(g))
    for b in (1, 3):
        f = oracle.simple_pk_exact({"k_abs": k_abs[b], "CL": CL[b]}, t)
        assert lp[b] == pytest.approx(oracle.loglik_normal(f, truth, 0.1), rel=1e-6)


def test_argument_errors_and_edge_cases(cuda_lib, rt, oracle):
    """Error behaviour of the boundary: wrong ids / shapes raise with a message, never compute garbage."""
    opt = pkg("optimization.petab_optimization")
    pb = rt.Problem("simple_pk")
    with pytest.raises(rt.ParvarB200Error, match="no settable id"):
        pb.set_columns(["k_abs", "nope"])
    with pytest.raises(rt.ParvarB200Error, match="no state"):
        pb.set_observables(["[y_gut]", "k_abs"])            # parameters are not observables
    with pytest.raises(rt.ParvarB200Error, match="strictly increasing"):
        pb.set_timepoints([0.0, 1.0, 1.0])
    with pytest.raises(rt.ParvarB200Error, match="t0 must be"):
        pb.set_timepoints([1.0, 2.0], t0=1.5)
    with pytest.raises(rt.ParvarB200Error, match="too many"):
        pb.set_columns(["k_abs", "CL", "Q", "Vgut", "Vperi", "Vcent", "y_gut", "y_cent", "y_peri"])
    with pytest.raises(rt.ParvarB200Error):
        rt.Problem("no_such_model")
    pb.set_columns(["k_abs"])
    with pytest.raises(rt.ParvarB200Error, match="timepoints"):
        pb.simulate_host(np.ones((1, 2)), 2)                  # grid not configured yet
    pb.set_timepoints(np.linspace(0, 10, 5))
    pb.set_observables(["y_cent"])
    with pytest.raises(rt.ParvarB200Error, match="data not set"):
        pb.logp_host(np.ones((1, 2)))
    with pytest.raises(ValueError):
        pb.simulate_host(np.ones((2, 2)), 2)                  # P has the wrong number of rows
    with pytest.raises(rt.ParvarB200Error, match="multiple of seg_len"):
        pb.set_data(np.ones((3, 5, 1, 1)), 1.0)
        pb.logp_host(np.ones((1, 5)), seg_len=2)
    with pytest.raises(rt.ParvarB200Error, match="sigma"):
        pb.set_data(np.ones((3, 5, 1, 1)), 0.0)
    # per-trajectory initial values and compartment sizes are settable columns as well (r.setValue on any id)
    t = np.linspace(0, 20, 9)
    pb.set_columns(["[y_gut]", "Vcent", "Q"])
    pb.set_timepoints(t)
    pb.set_observables(["y_gut", "y_cent", "y_peri"])
    P = np.array([[2.0, 0.5], [1.0, 3.0], [1.0, 0.2]])
    Y, nf = pb.simulate_host(P, 2)
    assert nf == 0
    for b in range(2):
        ref = oracle.simple_pk_exact({"y_gut": P[0, b], "Vcent": P[1, b], "Q": P[2, b]}, t)
        assert parity_violation(Y[b], ref) <= 1.0
    # missing measurements (NaN) are skipped: n = 0 entries contribute nothing, also under the log-normal model
    cols = [1, 2]
    truth = oracle.simple_pk_exact({}, t[1:])[:, cols]
    data = truth[None].repeat(3, axis=0) * 1.1
    data[0, 2, 0] = np.nan
    data[:, 5, 1] = np.nan
    for nm, sig in ((rt.NOISE_NORMAL, 0.1), (rt.NOISE_LOGNORMAL, 0.2)):
        pb2 = _problem(rt, "simple_pk", ["k_abs"], t[1:], ["y_cent", "y_peri"], rtol=1e-9, atol=1e-15)
        pb2.set_timepoints(t[1:], t0=0.0)
        pb2.set_data(opt.sufficient_statistics(data, nm)[..., None], sig, nm)
        lp, _, _, nf = pb2.logp_host(np.array([[0.8]]))
        f = oracle.simple_pk_exact({"k_abs": 0.8}, t[1:])[:, cols]
        ok = np.isfinite(data)
        ref = 0.0
        for i in range(3):
            m = ok[i]
            ref += (oracle.loglik_normal(f[m], data[i][m], sig) if nm == rt.NOISE_NORMAL else oracle.loglik_lognormal(f[m], data[i][m], sig))
        assert nf == 0 and lp[0] == pytest.approx(ref, rel=1e-6)


def test_sweep_extremes_timepoints_and_samples(cuda_lib, rt, oracle):
    """The corners of the reference's sweep (factories/__init__.py:12-20): 2 and 160 timepoints, 1 and 160 samples."""
    opt = pkg("optimization.petab_optimization")
    for T in (2, 160):
        t = np.linspace(0, 100, T)
        pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["y_cent", "y_gut", "y_peri"])
        th = np.stack([lognormal_draws(70, 160), lognormal_draws(71, 160)])
        Y, nf = pb.simulate_host(th, 160)
        assert nf == 0 and Y.shape == (160, T, 3)
        for b in (0, 1, 159):
            ref = oracle.simple_pk_exact({"k_abs": th[0, b], "CL": th[1, b]}, t)[:, [1, 0, 2]]
            assert parity_violation(Y[b], ref) <= 1.0
        # pooled likelihood over 160 individuals of one group at the sweep's largest cv (1.0; values may go negative)
        rs = np.random.RandomState(T)
        clean = oracle.simple_pk_exact({}, t)[:, [1, 0, 2]]
        data = clean[None] * (1 + 1.0 * rs.normal(size=(160,) + clean.shape))
        pb.set_data(opt.sufficient_statistics(data, rt.NOISE_NORMAL)[..., None], 1.0)
        lp, g, _, nf = pb.logp_host(np.array([[1.0], [1.0]]), grad=True)
        assert nf == 0 and lp[0] == pytest.approx(oracle.loglik_normal(clean, data, 1.0), rel=1e-6)


def test_objective_integrates_from_t0_and_poisons_failed_evaluations(cuda_lib, rt, oracle):
    """PEtab / AMICI semantics: the initial values hold at t = 0 even when the first measurement is later (PetabObjective
    like PetabObjectiveBatch); a trajectory that does not finish yields fval = inf and a NaN gradient (never a finite
    partial sum), with and without the Fisher information."""
    opt = pkg("optimization.petab_optimization")
    t = np.linspace(0, 30, 16)[3:]                       # measurements start at t = 6
    cols = [oracle.SIMPLE_PK.states.index(o) for o in ("y_cent", "y_peri")]
    truth = oracle.simple_pk_exact({}, t)[:, cols]
    y = _noisy_data(truth, 0.05, 11)[None]
    groups = [opt.GroupData("MALE", y), opt.GroupData("FEMALE", y * 1.1)]
    obj = opt.PetabObjective("simple_pk", ["k_abs", "CL"], groups, t, ["y_cent", "y_peri"], sigma=0.05, rtol=1e-9, atol=1e-13)
    x = np.array([0.9, 1.2, 1.1, 0.7])
    fval, g = obj(x, (0, 1))
    ref, ref_g = 0.0, np.zeros(4)
    for gi, gr in enumerate(groups):
        over = {"k_abs": x[gi], "CL": x[2 + gi]}
        f = oracle.simple_pk_exact(over, t)[:, cols]          # exact solution started at t = 0
        s = oracle.simple_pk_exact_sens(over, t)[:, cols, :]
        ref -= oracle.loglik_normal(f, gr.y[0], 0.05)
        gg = oracle.loglik_normal_grad(f, s, gr.y[0], 0.05)
        ref_g[gi], ref_g[2 + gi] = -gg[0], -gg[1]
    assert fval == pytest.approx(ref, rel=1e-6)
    np.testing.assert_allclose(g, ref_g, rtol=1e-6, atol=1e-6 * np.abs(ref_g).max())
    # failed evaluation: a step budget no trajectory can meet
    bad = opt.PetabObjective("simple_pk", ["k_abs", "CL"], groups, t, ["y_cent", "y_peri"], sigma=0.05, max_steps=5,
                             method=rt.METHOD_DOPRI5)
    fv, gb = bad(x, (0, 1))
    assert fv == np.inf and np.all(np.isnan(gb))
    lp, gr_, fim, _, nf = bad.problem.logp_fim_host(np.array([[0.9, 1.2], [1.1, 0.7]]))
    assert nf == 2 and np.all(np.isnan(lp)) and np.all(np.isnan(gr_)) and np.all(np.isnan(fim))


@pytest.mark.parametrize("B", [4099, 50000])
def test_locality_schedule_and_lockstep_warps(cuda_lib, rt, oracle, B):
    """pv_problem_set_locality: the device-side stable radix sort only changes the order in which trajectories are picked
    up - trajectories, logp and step counts are bit-identical with and without it (which also proves the index map is a
    permutation: a missing row would keep its poison value), for ragged batch sizes, duplicate, non-positive and NaN
    parameter values.  pv_problem_set_lockstep: one step size per warp - results inside the parity band against the exact
    solution, identical from run to run (deterministic work order), and a NaN row fails alone."""
    torch = _torch()
    T = 50
    t = np.linspace(0, 100, T)
    th = np.stack([lognormal_draws(71, B), lognormal_draws(72, B)])
    th[:, 100:140] = th[:, 100:101]                      # duplicates
    th[0, 7] = np.nan                                    # a row that cannot be integrated
    th[1, 11] = 0.0                                      # CL = 0: legal, key of a non-positive value
    pb = _problem(rt, "simple_pk", ["k_abs", "CL"], t, ["y_gut", "y_cent", "y_peri"], method=rt.METHOD_DOPRI5)
    truth = oracle.simple_pk_exact({}, t)
    pb.set_data(np.stack([np.ones_like(truth), truth, truth * truth])[..., None], 1.0)
    P = torch.from_numpy(th).cuda()

    def run(locality, lockstep):
        pb.set_locality(locality)
        pb.set_lockstep(lockstep)
        Y = torch.full((B, T, 3), -7.0, dtype=torch.float64, device="cuda")
        lp = torch.full((B,), -7.0, dtype=torch.float64, device="cuda")
        st = torch.full((B,), -7, dtype=torch.int32, device="cuda")
        steps = torch.full((2, B), -7, dtype=torch.int32, device="cuda")
        pb.simulate_logp(P, Y, lp, status=st, steps=steps)
        torch.cuda.synchronize()
        return Y.cpu().numpy(), lp.cpu().numpy(), st.cpu().numpy(), steps.cpu().numpy()

    Y0, lp0, st0, n0 = run(1, 1)
    Y1, lp1, st1, n1 = run(2, 1)
    ok = st0 == 0
    assert st0[7] != 0 and ok.sum() == B - 1
    assert np.array_equal(st0, st1) and np.array_equal(n0, n1)
    assert np.array_equal(Y0[ok], Y1[ok]) and np.array_equal(lp0[ok], lp1[ok])
    Y2, lp2, st2, n2 = run(2, 2)
    Y3, lp3, st3, n3 = run(2, 2)
    assert np.array_equal(st2, st0)                       # the NaN row fails alone, its 31 warp neighbours finish
    assert np.array_equal(Y2[ok], Y3[ok]) and np.array_equal(lp2[ok], lp3[ok]) and np.array_equal(n2, n3)
    for b in list(range(0, B, max(B // 40, 1))) + [11, 100, 139, B - 1]:
        ref = oracle.simple_pk_exact({"k_abs": th[0, b], "CL": th[1, b]}, t)
        assert parity_violation(Y2[b], ref) <= 1.0 and parity_violation(Y0[b], ref) <= 1.0
    np.testing.assert_allclose(lp2[ok], lp0[ok], rtol=1e-6)
    # neighbours in parameter space: a common step size costs few extra steps (the denser the batch, the fewer)
    assert n2[:, ok].sum() < (1.15 if B >= 50000 else 1.6) * n0[:, ok].sum()


@pytest.mark.parametrize("case", ["pk_all_T50", "pk_two_obs_T13", "pk_one_obs_T31", "chain_T9", "pk_reordered_T3"])
def test_lockstep_staged_trajectory_store(cuda_lib, rt, oracle, case):
    """The lockstep kernel stores trajectories through a per-warp shared-memory staging block (pv_sink STAGE: 24 doubles
    per row between flushes, the warp writes them out row-contiguously).  Every element of Y must be written exactly
    where the per-lane kernel writes it: full and partial flushes (T * n_obs not a multiple of 24), the compile-time
    all-states path and the run-time observable selection (n_obs = 1, 2, reordered), a grid whose first point is t0, a
    ragged last tile, with and without the locality schedule's index map, TRAJ and TRAJ|LOGP launches."""
    torch = _torch()
    model, cols, obs, T, sel = {
        "pk_all_T50": ("simple_pk", ["k_abs", "CL"], ["y_gut", "y_cent", "y_peri"], 50, [0, 1, 2]),
        "pk_two_obs_T13": ("simple_pk", ["k_abs", "CL"], ["y_cent", "y_peri"], 13, [1, 2]),
        "pk_one_obs_T31": ("simple_pk", ["k_abs", "CL"], ["y_peri"], 31, [2]),
        "chain_T9": ("simple_chain", ["k1"], ["S1", "S2"], 9, [0, 1]),
        "pk_reordered_T3": ("simple_pk", ["k_abs", "CL"], ["y_peri", "y_gut", "y_cent"], 3, [2, 0, 1]),
    }[case]
    B = 1000 + 13
    t = np.linspace(0, 40, T)
    th = np.stack([lognormal_draws(300 + i, B, 1.0, 0.5) for i in range(len(cols))])
    pb = _problem(rt, model, cols, t, obs, method=rt.METHOD_DOPRI5)
    P = torch.from_numpy(th).cuda()

    def run(locality, lockstep, with_logp):
        pb.set_locality(locality)
        pb.set_lockstep(lockstep)
        guard = 64
        buf = torch.full((B * T * len(obs) + 2 * guard,), -7.0, dtype=torch.float64, device="cuda")
        Y = buf[guard:guard + B * T * len(obs)].view(B, T, len(obs))
        st = torch.full((B,), -7, dtype=torch.int32, device="cuda")
        if with_logp:
            lp = torch.empty((B,), dtype=torch.float64, device="cuda")
            pb.simulate_logp(P, Y, lp, status=st)
        else:
            pb.simulate(P, Y, status=st)
        torch.cuda.synchronize()
        assert int((st != 0).sum()) == 0
        g = buf.cpu().numpy()
        assert np.all(g[:guard] == -7.0) and np.all(g[-guard:] == -7.0)      # nothing outside the rows
        return Y.cpu().numpy()

    Y_lane = run(1, 1, False)                                                  # per-lane kernel, direct stores
    if model == "simple_pk":
        ref = np.stack([oracle.simple_pk_exact({"k_abs": th[0, b], "CL": th[1, b]}, t)[:, sel] for b in range(B)])
    else:
        ref = np.stack([oracle.simple_chain_exact(th[0, b], t)[:, sel] for b in range(B)])
    assert parity_violation(Y_lane, ref) <= 1.0
    data = ref[0]
    pb.set_data(np.stack([np.ones_like(data), data, data * data])[..., None], 1.0)
    for locality in (1, 2):
        Y_a = run(locality, 2, False)
        Y_b = run(locality, 2, True)
        assert not np.any(Y_a == -7.0)                                         # every element written
        assert parity_violation(Y_a, ref) <= 1.0
        assert np.array_equal(Y_a, Y_b)                                        # TRAJ and TRAJ|LOGP: the same trajectories
        assert parity_violation(Y_a, Y_lane, rtol=1e-6, atol=1e-9) <= 1.0


@pytest.mark.parametrize("noise", ["normal", "lognormal"])
@pytest.mark.parametrize("model", ["simple_pk", "simple_chain"])
def test_group_lanes_kernel_matches_lane_kernel_and_oracle(cuda_lib, rt, oracle, model, noise):
    """pv_problem_set_group_lanes: the small-batch DOPRI5 logp + gradient kernel that spreads a trajectory over n_sens
    lanes (pv_dopri5_pairs.cuh) against the one-trajectory-per-lane kernel and the exact solution: ragged batch (not a
    multiple of the trajectories per warp), per-trajectory data sets and a pooled one, both noise models, a NaN row that
    fails alone, sensitivity columns in a non-compiled order."""
    opt = pkg("optimization.petab_optimization")
    nm = rt.NOISE_NORMAL if noise == "normal" else rt.NOISE_LOGNORMAL
    B = 203
    if model == "simple_pk":
        cols, obs = ["CL", "k_abs"], ["y_cent", "y_peri"]
        t = np.linspace(0, 30, 13)[1:]
        # log-normal noise: log f amplifies the (absolute-tolerance sized) differences of outputs that have decayed to
        # 1e-10; narrow draws keep every output far above atol, as in test_simple_pk_logp_and_grad
        sd = 1.0 if noise == "normal" else 0.3
        th = np.stack([lognormal_draws(91, B, 1.0, sd), lognormal_draws(92, B, 1.0, sd)])
        exact = lambda b: (oracle.simple_pk_exact({"CL": th[0, b], "k_abs": th[1, b]}, t)[:, [1, 2]],           # noqa: E731
                           oracle.simple_pk_exact_sens({"CL": th[0, b], "k_abs": th[1, b]}, t, sens=("CL", "k_abs"))[:, [1, 2], :])
        truth = oracle.simple_pk_exact({}, t)[:, [1, 2]]
    else:
        cols, obs = ["k1"], ["S1", "S2"]
        t = np.linspace(0, 8, 9)[1:]
        th = lognormal_draws(93, B, 1.0, 1.0 if noise == "normal" else 0.3)[None, :]
        exact = lambda b: (oracle.simple_chain_exact(th[0, b], t), oracle.simple_chain_exact_sens(th[0, b], t))  # noqa: E731
        truth = oracle.simple_chain_exact(1.0, t)
    th[0, 17] = np.nan
    rs = np.random.RandomState(3)
    for n_data in (1, 7):
        data = np.abs(truth[None] * (1 + 0.05 * rs.normal(size=(n_data,) + truth.shape))) + 1e-9
        stats = np.stack([opt.sufficient_statistics(data[i:i + 1], nm) for i in range(n_data)], axis=-1)
        sigma = 0.2 if nm == rt.NOISE_LOGNORMAL else 0.05
        pb = _problem(rt, model, cols, t, obs, rtol=1e-9, atol=1e-14, method=rt.METHOD_DOPRI5)
        pb.set_timepoints(t, t0=0.0)
        pb.set_data(stats, sigma, nm)
        out = {}
        for mode in (1, 2):
            pb.set_group_lanes(mode)
            n0 = rt.launch_count()
            lp, g, _, nf = pb.logp_host(th, grad=True)
            assert nf == 1 and rt.launch_count() > n0
            out[mode] = (lp, g)
        ok = np.arange(B) != 17
        assert np.isnan(out[2][0][17]) and np.all(np.isnan(out[2][1][:, 17])) and np.all(np.isfinite(out[2][0][ok]))
        # two step sequences of the same method at rtol 1e-9; the log-normal likelihood multiplies the relative error of f by
        # residual / sigma^2 (a few hundred here)
        np.testing.assert_allclose(out[2][0][ok], out[1][0][ok], rtol=1e-8 if nm == rt.NOISE_NORMAL else 1e-6)
        np.testing.assert_allclose(out[2][1][:, ok], out[1][1][:, ok], rtol=1e-6, atol=1e-6 * np.abs(out[1][1][:, ok]).max())
        rows = [pb.sens_ids.index(c) for c in cols]
        for b in (0, 9, 10, 11, 200, 202):
            f, s = exact(b)
            d = data[b % n_data]
            if nm == rt.NOISE_NORMAL:
                ref_lp, ref_g = oracle.loglik_normal(f, d, sigma), oracle.loglik_normal_grad(f, s, d, sigma)
            else:
                ref_lp, ref_g = oracle.loglik_lognormal(f, d, sigma), oracle.loglik_lognormal_grad(f, s, d, sigma)
            assert out[2][0][b] == pytest.approx(ref_lp, rel=1e-6)
            np.testing.assert_allclose(out[2][1][rows, b], ref_g, rtol=1e-6, atol=1e-6 * np.abs(ref_g).max())


def test_unregistered_sbml_model_and_other_sensitivity_sets_end_to_end(cuda_lib, rt, oracle, tmp_path):
    """Run-time model compile (petab_factory.py:183-190: the reference loads ANY SBML file): a perturbed copy of
    simple_pk.xml - extra saturable elimination, two new parameters - goes SBML -> generated CUDA -> nvcc -> cached library
    -> batched simulation, checked against scipy on the same ODE; its gradient with respect to a sensitivity set chosen at
    run time against finite differences of that oracle.  And a registry model asked for a parameter that is NOT among its
    compiled sensitivity parameters (simple_pk, Q) is rebuilt for that set instead of raising."""
    import pandas as pd
    from scipy.integrate import solve_ivp
    from test_codegen import perturbed_simple_pk
    pf = pkg("experiments.petab_factory")
    ex = pkg("experiments.experiment")
    opt = pkg("optimization.petab_optimization")
    path = perturbed_simple_pk(tmp_path)
    t = np.linspace(0, 30, 16)

    def solve(k_abs, k_el, CL=1.0, Q=1.0, Km=0.05):
        def f(_, y):
            return [-k_abs * y[0], k_abs * y[0] - (CL + Q) * y[1] + Q * y[2], Q * y[1] - Q * y[2] - k_el * y[2] / (Km + y[2])]
        return solve_ivp(f, (0, 30), [1.0, 0.0, 0.0], method="LSODA", t_eval=t, rtol=1e-12, atol=1e-14).y.T

    # ---- B1: the simulator mirror on the unknown file ----------------------------------------------------------------
    sim = pf.ODESampleSimulator(path)
    assert sim.model_name.startswith("pk_saturable@") and "k_el" in sim.ids and "[y_peri]" in sim.ids
    n = 9
    pars = pd.DataFrame({"k_abs": lognormal_draws(5, n), "k_el": lognormal_draws(6, n, 0.3, 0.1)})
    st = pf.SimulationSettings(start=0.0, end=30.0, steps=15, dosage=None, noise=ex.Noise(add_noise=False, cv=0.0),
                               observables=[ex.Observable(id=o) for o in ("y_gut", "y_cent", "y_peri")])
    ds = sim.simulate_samples(pars, st)
    got = np.stack([ds[f"[{o}]"].values for o in ("y_gut", "y_cent", "y_peri")], axis=2)
    for i in range(n):
        assert parity_violation(got[i], solve(pars["k_abs"][i], pars["k_el"][i])) <= 1.0
    # ---- B2: gradient with respect to (k_abs, k_el), a set chosen now -------------------------------------------------
    name = pf.resolve_model(path, sens=["k_abs", "k_el"])
    pb = _problem(rt, name, ["k_abs", "k_el"], t, ["y_cent", "y_peri"], rtol=1e-9, atol=1e-13, method=rt.METHOD_DOPRI5)
    assert pb.sens_ids == ["k_abs", "k_el"]
    data = _noisy_data(solve(1.0, 0.3)[:, 1:], 0.05, 4)[None]
    sigma = 0.02
    pb.set_data(opt.sufficient_statistics(data, rt.NOISE_NORMAL)[..., None], sigma, rt.NOISE_NORMAL)
    th = np.stack([pars["k_abs"].to_numpy(), pars["k_el"].to_numpy()])
    for mode in (1, 2):                                   # lane kernel / group-lanes kernel (rhs not linear in c: select path)
        pb.set_group_lanes(mode)
        lp, g, _, nf = pb.logp_host(th, grad=True)
        assert nf == 0
        for i in (0, 4, 8):
            ka, ke = th[:, i]
            ll = lambda a, e: oracle.loglik_normal(solve(a, e)[:, 1:], data[0], sigma)     # noqa: E731
            assert lp[i] == pytest.approx(ll(ka, ke), rel=1e-6)
            fd = np.array([(ll(ka * (1 + 1e-5), ke) - ll(ka * (1 - 1e-5), ke)) / (2e-5 * ka),
                           (ll(ka, ke * (1 + 1e-5)) - ll(ka, ke * (1 - 1e-5))) / (2e-5 * ke)])
            np.testing.assert_allclose(g[:, i], fd, rtol=2e-4, atol=2e-4 * np.abs(fd).max())
    # ---- a registry model, another sensitivity set ---------------------------------------------------------------------
    cols = [oracle.SIMPLE_PK.states.index(o) for o in ("y_cent", "y_peri")]
    y = _noisy_data(oracle.simple_pk_exact({}, t)[:, cols], 0.05, 8)[None]
    obj = opt.PetabObjective("simple_pk", ["Q"], [opt.GroupData("MALE", y)], t, ["y_cent", "y_peri"], sigma=0.05, rtol=1e-9, atol=1e-13)
    assert obj.problem.sens_ids == ["Q"] and obj.problem.model.startswith("simple_pk@")
    fval, grad = obj(np.array([1.3]), (0, 1))
    f = oracle.simple_pk_exact({"Q": 1.3}, t)[:, cols]
    s = oracle.simple_pk_exact_sens({"Q": 1.3}, t, sens=("Q",))[:, cols, :]
    assert fval == pytest.approx(-oracle.loglik_normal(f, y[0], 0.05), rel=1e-6)
    np.testing.assert_allclose(grad, -oracle.loglik_normal_grad(f, s, y[0], 0.05), rtol=1e-6)
