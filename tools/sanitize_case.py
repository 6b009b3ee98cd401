#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck): every kernel mode of two models on tiny batches."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rt = importlib.import_module("parameter-variability_b200.runtime")
t = np.linspace(0, 20, 7)
for model, cols, obs, dose in (("simple_pk", ["k_abs", "CL"], ["y_gut", "y_cent", "y_peri"], {}),
                               ("simple_pk", ["k_abs", "CL"], ["y_peri", "y_cent"], {}),
                               ("icg_body_flat", ["BW", "LI__ICGIM_Vmax"], ["Cre_plasma_icg", "Cgi_plasma_icg"], {"IVDOSE_icg": 10.0})):
    pb = rt.Problem(model)
    for k, v in dose.items():
        pb.set_value(k, v)
    pb.set_columns(cols)
    pb.set_timepoints(t)
    pb.set_observables(obs)
    B = 77
    rs = np.random.RandomState(0)
    base = np.array([pb.theta_default[pb.theta_ids.index(c)] for c in cols])
    P = base[:, None] * rs.lognormal(0, 0.2, size=(len(cols), B))
    if model == "simple_pk":
        P[0, 3] = 1e7                                      # one stiff row -> exercises the RODAS4 re-run pass
    status = np.empty(B, np.int32)
    steps = np.empty((2, B), np.int32)
    Y, nf = pb.simulate_host(P, B, status=status, steps=steps)
    assert nf == 0, (model, status)
    for n_data in (1, 7):
        stats = np.stack([np.ones((len(t), len(obs), n_data)), np.full((len(t), len(obs), n_data), 1e-3),
                          np.full((len(t), len(obs), n_data), 1e-6)])
        pb.set_data(stats, 0.1)
        lp, g, seg, nf = pb.logp_host(P[:, :70], seg_len=7, grad=True)
        lp2, g2, f2, seg2, nf2 = pb.logp_fim_host(P[:, :70], seg_len=7)
        lpo = np.empty(B)
        Y2, nf3 = pb.simulate_host(P, B, logp=lpo)
        assert nf == 0 and nf2 == 0 and nf3 == 0 and np.all(np.isfinite(lp)) and np.allclose(lp, lp2, rtol=1e-6)
    if model == "simple_pk":
        # round-2 kernels: lockstep warps with staged trajectory stores (full and partial flushes, ragged last tile), the
        # small-batch sensitivity kernel (one trajectory per group of n_sens lanes), both with explicit DOPRI5
        pb.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-7, atol=1e-10)
        P[0, 3] = 1.3
        pb.set_lockstep(2)
        for locality in (1, 2):
            pb.set_locality(locality)
            lpo = np.empty(B)
            Y3, nf4 = pb.simulate_host(P, B, logp=lpo)
            Y4, nf5 = pb.simulate_host(P, B)
            assert nf4 == 0 and nf5 == 0 and np.array_equal(Y3, Y4) and np.all(np.isfinite(Y3))
        pb.set_lockstep(0)
        pb.set_locality(0)
        pb.set_group_lanes(2)
        lp3, g3, seg3, nf6 = pb.logp_host(P[:, :70], seg_len=7, grad=True)
        pb.set_group_lanes(1)
        lp4, g4, seg4, nf7 = pb.logp_host(P[:, :70], seg_len=7, grad=True)
        assert nf6 == 0 and nf7 == 0 and np.allclose(lp3, lp4, rtol=1e-6) and np.allclose(g3, g4, rtol=1e-5, atol=1e-8)
    pb.close()

# device-resident HMC on a tiny hierarchical posterior (pv_hmc_* kernels around pv_hier_logp_grad)
import torch  # noqa: E402
po = importlib.import_module("parameter-variability_b200.optimization.petab_optimization")
hm = importlib.import_module("parameter-variability_b200.optimization.hmc")
n_ind, C, Pn = 5, 2, 2
pb = rt.Problem("simple_pk")
pb.set_solver(rtol=1e-7, atol=1e-10)
pb.set_columns(["k_abs", "CL"])
pb.set_timepoints(t[1:], t0=0.0)
pb.set_observables(["y_gut", "y_cent", "y_peri"])
Y1 = torch.empty((1, len(t) - 1, 3), dtype=torch.float64, device="cuda")
pb.simulate(torch.ones((2, 1), dtype=torch.float64, device="cuda"), Y1)
torch.cuda.synchronize()
y = Y1.cpu().numpy() * (1 + 0.05 * np.random.RandomState(1).normal(size=(n_ind, len(t) - 1, 3)))
pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), 0.05)
Hd = po.HierarchicalObjectiveDevice(pb, C, n_ind, mu_prior=[(0.0, 1.0)] * Pn, omega_prior=[0.5] * Pn)
x0 = np.concatenate([0.05 * np.random.RandomState(2).normal(size=(C, n_ind * Pn)), np.zeros((C, Pn)), np.log(0.3) * np.ones((C, Pn))], axis=1)
res = hm.hmc_device(Hd, x0, n_samples=3, n_warmup=14, n_leapfrog=3, step_size=0.02, seed=3)
assert np.all(np.isfinite(res.samples)) and np.all(np.isfinite(res.logp))
nt = importlib.import_module("parameter-variability_b200.optimization.nuts")
rn = nt.nuts_device(Hd, x0, n_samples=2, n_warmup=14, step_size=0.02, max_depth=3, seed=3)
assert np.all(np.isfinite(rn.samples)) and rn.tree_depth.max() <= 3
pb.close()
print("sanitize case ok")
