"""Committed golden vectors (tests/golden/, made by tools/make_golden.py from the oracle):
  - not-gpu: the oracle (Python and C restatements) still reproduces them  -> pins the oracle against drift;
  - gpu:     the CUDA path reproduces them through the C ABI."""
from pathlib import Path

import numpy as np
import pytest

from conftest import parity_violation, pkg

G = Path(__file__).resolve().parent / "golden"


def test_oracle_reproduces_golden(oracle):
    g = np.load(G / "simple_pk_c2.npz")
    for b in (0, 17, 63):
        over = {"k_abs": g["k_abs"][b], "CL": g["CL"][b]}
        np.testing.assert_allclose(oracle.simple_pk_exact(over, g["t"]), g["Y"][:, :, b], rtol=1e-12, atol=1e-300)
        assert oracle.loglik_normal(g["Y"][:, :, b], g["data"], float(g["sigma"])) == pytest.approx(g["logp"][b], rel=1e-13)
    g = np.load(G / "icg_body_flat_c4.npz")
    Y = oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0, "BW": g["BW"][5], "LI__ICGIM_Vmax": g["Vmax"][5]}, g["t"])
    assert parity_violation(Y, g["Y"][:, :, 5], rtol=1e-8, atol=1e-12) <= 1.0
    g = np.load(G / "sampling_simple_pk.npz")
    s = oracle.create_samples_parameters({"MALE": {"CL": (20, 1.0, 1.0), "k_abs": (20, 1.0, 0.5)},
                                          "FEMALE": {"CL": (20, 1.0, 0.5), "k_abs": (20, 1.0, 1.0)}})
    np.testing.assert_array_equal(s["FEMALE"]["k_abs"], g["k_abs_FEMALE"])
    np.testing.assert_array_equal(oracle.add_noise(g["y"], 0.05), g["noisy"])


def test_host_mirror_sampling_matches_golden():
    """create_samples_parameters of the host mirror = the reference's stream, bit for bit."""
    pf = pkg("experiments.petab_factory")
    g = np.load(G / "sampling_simple_pk.npz")
    L = pf.LognormParameters
    s = pf.create_samples_parameters({
        pf.Category.MALE: {pf.PKPDParameters.CL: L(n=20, sigma=1.0, mu=1.0), pf.PKPDParameters.k_abs: L(n=20, sigma=1.0, mu=0.5)},
        pf.Category.FEMALE: {pf.PKPDParameters.CL: L(n=20, sigma=1.0, mu=0.5), pf.PKPDParameters.k_abs: L(n=20, sigma=1.0, mu=1.0)}})
    np.testing.assert_array_equal(s[pf.Category.MALE][pf.PKPDParameters.CL], g["CL_MALE"])
    np.testing.assert_array_equal(s[pf.Category.FEMALE][pf.PKPDParameters.k_abs], g["k_abs_FEMALE"])


def test_c_oracle_matches_golden():
    """The C restatement (the bench's CPU baseline) integrates the same systems: variable-order BDF at tight
    tolerance vs the golden trajectories, and at the reference's 1e-6/1e-6 within the accuracy those settings give."""
    import oracle_c
    g = np.load(G / "simple_pk_c2.npz")
    obs = ["y_gut", "y_cent", "y_peri"]
    Y, lp, nf, st = oracle_c.simulate_batch("simple_pk", {"k_abs": g["k_abs"], "CL": g["CL"]}, g["t"], obs,
                                            rtol=1e-9, atol=1e-12, data=g["data"], sigma=float(g["sigma"]))
    assert nf == 0 and parity_violation(Y, g["Y"]) <= 1.0
    np.testing.assert_allclose(lp, g["logp"], rtol=1e-6)
    Y6, _, nf, _ = oracle_c.simulate_batch("simple_pk", {"k_abs": g["k_abs"], "CL": g["CL"]}, g["t"], obs, rtol=1e-6, atol=1e-6)
    assert nf == 0 and np.max(np.abs(Y6 - g["Y"])) < 2e-5          # quirk Q9: the reference's own data accuracy
    g = np.load(G / "icg_body_flat_c4.npz")
    states = list(g["states"])
    Y, _, nf, _ = oracle_c.simulate_batch("icg_body_flat", {"BW": g["BW"], "LI__ICGIM_Vmax": g["Vmax"]}, g["t"], states,
                                          overrides={"IVDOSE_icg": 10.0}, rtol=1e-9, atol=1e-13)
    assert nf == 0 and parity_violation(Y, g["Y"]) <= 1.0
    g = np.load(G / "simple_chain.npz")
    Y, _, nf, _ = oracle_c.simulate_batch("simple_chain", {"k1": g["k1"]}, g["t"], ["S1", "S2"], rtol=1e-9, atol=1e-12)
    assert nf == 0 and parity_violation(Y, g["Y"]) <= 1.0


def test_c_oracle_rhs_matches_python_oracle(oracle):
    import ctypes
    import oracle_c
    lib = oracle_c.load()
    rs = np.random.RandomState(0)
    for name in ("simple_chain", "simple_pk", "icg_body_flat"):
        m, states, params, defaults = oracle_c.model_info(name)
        om = oracle.MODELS[name]
        assert states == om.states
        p = np.array(defaults) * rs.lognormal(0, 0.2, len(defaults))
        y = rs.rand(len(states)) * 1e-2
        f = np.zeros(len(states))
        dp = ctypes.POINTER(ctypes.c_double)
        lib.pvo_rhs(m, y.ctypes.data_as(dp), p.ctypes.data_as(dp), f.ctypes.data_as(dp))
        ref = np.array(om.rhs(0.0, y, dict(zip(params, p))), dtype=float)
        np.testing.assert_allclose(f, ref, rtol=1e-12, atol=1e-300)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cuda_reproduces_golden_simple_pk(cuda_lib, rt):
    g = np.load(G / "simple_pk_c2.npz")
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-7, atol=1e-10)
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(g["t"])
    pb.set_observables(list(g["states"]))
    th = np.stack([g["k_abs"], g["CL"]])
    data = g["data"]
    pb.set_data(np.stack([np.ones_like(data), data, data * data])[..., None], float(g["sigma"]))
    lp_out = np.empty(len(g["k_abs"]))
    Y, nf = pb.simulate_host(th, th.shape[1], logp=lp_out)
    assert nf == 0 and parity_violation(Y, np.moveaxis(g["Y"], -1, 0)) <= 1.0       # golden files are [T, obs, B]
    np.testing.assert_allclose(lp_out, g["logp"], rtol=1e-6)
    pb.set_solver(rtol=1e-9, atol=1e-13)
    lp, gr, _, nf = pb.logp_host(th, grad=True)
    np.testing.assert_allclose(lp, g["logp"], rtol=1e-6)
    np.testing.assert_allclose(gr, g["grad"], rtol=1e-6, atol=1e-6 * np.abs(g["grad"]).max())


@pytest.mark.gpu
def test_cuda_reproduces_golden_chain_and_icg(cuda_lib, rt):
    import torch
    g = np.load(G / "simple_chain.npz")
    pb = rt.Problem("simple_chain")
    pb.set_columns(["k1"])
    pb.set_timepoints(g["t"])
    pb.set_observables(["S1", "S2"])
    Y, nf = pb.simulate_host(g["k1"][None, :], len(g["k1"]))
    assert nf == 0 and parity_violation(Y, np.moveaxis(g["Y"], -1, 0)) <= 1.0
    B, T = len(g["k1"]), len(g["t"])
    pb.set_solver(rtol=1e-9, atol=1e-12)
    P = torch.from_numpy(g["k1"][None, :]).cuda()
    Yd = torch.empty((B, T, 2), dtype=torch.float64, device="cuda")
    S = torch.empty((B, T, 2, 1), dtype=torch.float64, device="cuda")
    pb.simulate_sens(P, Yd, S)
    torch.cuda.synchronize()
    assert parity_violation(S[..., 0].cpu().numpy(), np.moveaxis(g["S"][:, :, 0, :], -1, 0)) <= 1.0

    g = np.load(G / "icg_body_flat_c4.npz")
    pb = rt.Problem("icg_body_flat")
    pb.set_value("IVDOSE_icg", 10.0)
    pb.set_columns(["BW", "LI__ICGIM_Vmax"])
    pb.set_timepoints(g["t"])
    pb.set_observables(list(g["states"]))
    th = np.stack([g["BW"], g["Vmax"]])
    Y, nf = pb.simulate_host(th, th.shape[1])
    assert nf == 0 and parity_violation(Y, np.moveaxis(g["Y"], -1, 0)) <= 1.0
    pb.set_observables(list(g["obs"]))
    pb.set_solver(rtol=1e-9, atol=1e-14)
    data = g["data"]
    pb.set_data(np.stack([np.ones_like(data), data, data * data])[..., None], float(g["sigma"]))
    lp, gr, _, nf = pb.logp_host(th, grad=True)
    assert nf == 0
    np.testing.assert_allclose(lp, g["logp"], rtol=1e-6)
    np.testing.assert_allclose(gr[:, :4], g["grad"], rtol=1e-5)
