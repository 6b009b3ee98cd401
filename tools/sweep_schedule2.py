#!/usr/bin/env python
"""Refill-threshold sweep for the gradient (DOPRI5 + sensitivities) and the stiff (RODAS4) kernels."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

rt = importlib.import_module("parameter-variability_b200.runtime")


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B, T = 256_000, 20
t = np.linspace(0, 100, T)
rs = np.random.RandomState(0)
th = rs.lognormal(-0.3466, 0.8326, size=(2, B))
P = torch.from_numpy(th).cuda()
lp = torch.empty(B, dtype=torch.float64, device="cuda")
g = torch.empty((2, B), dtype=torch.float64, device="cuda")
pb = rt.Problem("simple_pk")
pb.set_solver(rtol=1e-8, atol=1e-11)
pb.set_columns(["k_abs", "CL"])
pb.set_timepoints(t)
pb.set_observables(["y_gut", "y_cent", "y_peri"])
pb.set_data(np.ones((3, T, 3, 1)), 1.0)
for rm in (1, 3, 6, 10, 16, 32):
    pb.set_schedule(rm)
    print(f"simple_pk logp+grad  refill_min {rm:2d}: {timeit(lambda: pb.logp_grad(P, lp, g)):.3f} ms", flush=True)

B = 128_000
th = np.stack([rs.lognormal(np.log(75), 0.13, B), rs.lognormal(np.log(0.037), 0.27, B)])
P = torch.from_numpy(th).cuda()
Y = torch.empty((B, T, 2), dtype=torch.float64, device="cuda")
lp = torch.empty(B, dtype=torch.float64, device="cuda")
g = torch.empty((2, B), dtype=torch.float64, device="cuda")
pi = rt.Problem("icg_body_flat")
pi.set_solver(rtol=1e-7, atol=1e-11)
pi.set_value("IVDOSE_icg", 10.0)
pi.set_columns(["BW", "LI__ICGIM_Vmax"])
pi.set_timepoints(t)
pi.set_observables(["Cre_plasma_icg", "Cgi_plasma_icg"])
pi.set_data(np.ones((3, T, 2, 1)) * 1e-3, 1e-4)
for rm in (1, 3, 6, 10, 16, 32):
    pi.set_schedule(rm)
    print(f"icg forward          refill_min {rm:2d}: {timeit(lambda: pi.simulate(P, Y)):.3f} ms", flush=True)
for rm in (1, 6, 16, 32):
    pi.set_schedule(rm)
    print(f"icg logp+grad        refill_min {rm:2d}: {timeit(lambda: pi.logp_grad(P, lp, g), 3):.3f} ms", flush=True)
