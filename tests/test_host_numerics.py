"""The integrator templates (pv_dopri5.cuh, pv_rosenbrock.cuh) + generated model code, compiled as host C++
(tests/host/host_harness.cpp), against the oracle.  This checks the numerics of the exact source the CUDA
kernels are built from in the GPU-less container; the GPU parity tests repeat it through the C ABI."""
import numpy as np
import pytest

from conftest import lognormal_draws, parity_violation
from host import harness

DOPRI5, RODAS4 = 1, 2


def test_dopri5_simple_pk_parity_over_parameter_cloud(oracle):
    t = np.linspace(0, 100, 50)
    k_abs, CL = lognormal_draws(1, 40), lognormal_draws(2, 40)
    worst, steps = 0.0, []
    for a, c in zip(k_abs, CL):
        rc, Y, st = harness.simulate("simple_pk", {"k_abs": a, "CL": c}, t, 1e-7, 1e-10, DOPRI5)
        assert rc == 0
        worst = max(worst, parity_violation(Y, oracle.simple_pk_exact({"k_abs": a, "CL": c}, t)))
        steps.append(sum(st))
    assert worst <= 1.0
    assert 50 < np.mean(steps) < 2000


def test_rodas4_icg_parity(oracle):
    t = np.linspace(0, 100, 21)
    for bw, vm in [(75.0, 0.0369598840327503), (58.0, 0.02), (101.0, 0.06)]:
        over = {"IVDOSE_icg": 10.0, "BW": bw, "LI__ICGIM_Vmax": vm}
        rc, Y, st = harness.simulate("icg_body_flat", over, t, 1e-7, 1e-10, RODAS4)
        assert rc == 0 and st[0] < 2000
        m = harness.meta("icg_body_flat")
        ref = oracle.simulate("icg_body_flat", over, t)[:, [oracle.ICG_STATES.index(s) for s in m["states"]]]
        assert parity_violation(Y, ref) <= 1.0


def test_rodas4_on_nonstiff_models(oracle):
    t = np.linspace(0, 100, 21)
    rc, Y, st = harness.simulate("simple_pk", {}, t, 1e-7, 1e-10, RODAS4)
    assert rc == 0 and parity_violation(Y, oracle.simple_pk_exact({}, t)) <= 1.0
    rc, Y, st = harness.simulate("simple_chain", {"k1": 2.5}, np.linspace(0, 10, 11), 1e-7, 1e-10, RODAS4)
    assert rc == 0 and parity_violation(Y, oracle.simple_chain_exact(2.5, np.linspace(0, 10, 11))) <= 1.0


@pytest.mark.parametrize("method", [DOPRI5, RODAS4])
def test_forward_sensitivities_simple_pk(method, oracle):
    t = np.linspace(0, 100, 21)
    over = {"k_abs": 0.7, "CL": 1.6}
    rc, Z, st = harness.simulate("simple_pk", over, t, 1e-9, 1e-12, method, sens=True)
    assert rc == 0
    S = oracle.simple_pk_exact_sens(over, t)
    assert parity_violation(Z[:, :3], oracle.simple_pk_exact(over, t)) <= 1.0
    assert parity_violation(Z[:, 3:6], S[:, :, 0]) <= 1.0
    assert parity_violation(Z[:, 6:9], S[:, :, 1]) <= 1.0


def test_rodas4_sensitivities_icg(oracle):
    """Rosenbrock on the augmented system with the exact block Jacobian (second-order term from the generator)."""
    t = np.linspace(0, 50, 11)
    over = {"IVDOSE_icg": 10.0, "BW": 82.0}
    rc, Z, st = harness.simulate("icg_body_flat", over, t, 1e-9, 1e-14, RODAS4, sens=True)
    assert rc == 0
    m = harness.meta("icg_body_flat")
    perm = [oracle.ICG_STATES.index(s) for s in m["states"]]
    Y, S = oracle.simulate_sens("icg_body_flat", over, t, ["BW", "LI__ICGIM_Vmax"])
    assert parity_violation(Z[:, :12], Y[:, perm]) <= 1.0
    for k in range(2):
        ref = S[:, perm, k]
        got = Z[:, 12 * (1 + k):12 * (2 + k)]
        scale = np.abs(ref).max(axis=0) + 1e-6 * np.abs(ref).max()
        assert np.max(np.abs(got - ref) / scale) < 1e-6


def test_rodas4_convergence_order(oracle):
    """The RODAS4 coefficient set was written from memory of Hairer & Wanner's rodas.f, so its order is verified
    numerically on the (nonlinear, stiff) icg model: halving a fixed step must cut the error ~16x (order 4)."""
    over = {"IVDOSE_icg": 10.0}
    m = harness.meta("icg_body_flat")
    perm = [oracle.ICG_STATES.index(s) for s in m["states"]]
    ref = oracle.simulate("icg_body_flat", over, np.array([0.0, 2.0]), rtol=1e-13, atol=1e-18)[-1, perm]
    errs = []
    for n in (50, 100, 200):
        y = harness.rodas4_fixed("icg_body_flat", over, 2.0 / n, n)
        errs.append(np.max(np.abs(y - ref) / (np.abs(ref) + 1e-9)))
    orders = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(orders > 3.7) and errs[-1] < 1e-6, (errs, orders)


def test_status_codes():
    t = np.linspace(0, 100, 5)
    rc, Y, st = harness.simulate("simple_pk", {}, t, 1e-7, 1e-10, DOPRI5, max_steps=5)
    assert rc == 1
    rc, Y, st = harness.simulate("simple_pk", {"k_abs": float("nan")}, t, 1e-7, 1e-10, DOPRI5)
    assert rc != 0


def test_dopri5_stiffness_detection_hands_over_early_and_only_when_it_pays():
    """PV_METHOD_AUTO's explicit first pass (pv_dopri5_lane::stiff_after): rate constants near the PEtab upper bound end
    the explicit attempt after ~100 steps instead of the 4000-step budget; the benchmark's log-normal cloud - stability
    limited in its flat tail, but cheap - is never handed over."""
    t = np.linspace(0, 100, 50)
    for k, cl in ((1e5, 1.0), (1.0, 1e6), (1e8, 1e8)):
        rc0, _, st0 = harness.simulate("simple_pk", {"k_abs": k, "CL": cl}, t, 1e-7, 1e-10, max_steps=4000)
        rc1, _, st1 = harness.simulate("simple_pk", {"k_abs": k, "CL": cl}, t, 1e-7, 1e-10, max_steps=4000, stiff_after=64)
        assert rc0 == 1 and sum(st0) == 4000          # PV_ERR_MAX_STEPS after the whole budget
        assert rc1 == 4 and sum(st1) < 600            # PV_ERR_STIFF
    rs = np.random.RandomState(1234)
    th = rs.lognormal(-0.3465735902799726, 0.8325546111576977, size=(2, 300))
    for b in range(300):
        over = {"k_abs": th[0, b], "CL": th[1, b]}
        rc0, Y0, st0 = harness.simulate("simple_pk", over, t, 1e-7, 1e-10)
        rc1, Y1, st1 = harness.simulate("simple_pk", over, t, 1e-7, 1e-10, stiff_after=64)
        assert rc0 == 0 and rc1 == 0 and st0 == st1 and np.array_equal(Y0, Y1)
