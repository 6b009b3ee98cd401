#!/usr/bin/env python
"""Sweep the warp-refill threshold of the batch kernel on the C2 workload (tuning aid; prints ms per 1e6 trajectories)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

rt = importlib.import_module("parameter-variability_b200.runtime")
B, T = 1_000_000, 50
np.random.seed(1234)
th = np.stack([np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B) for _ in range(2)])
P = torch.from_numpy(th).cuda()
Y = torch.empty((B, T, 3), dtype=torch.float64, device="cuda")
lp = torch.empty(B, dtype=torch.float64, device="cuda")
t = np.linspace(0, 100, T)
pb = rt.Problem("simple_pk")
pb.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-7, atol=1e-10)
pb.set_columns(["k_abs", "CL"])
pb.set_timepoints(t)
pb.set_observables(["y_gut", "y_cent", "y_peri"])
d = np.ones((T, 3))
pb.set_data(np.stack([d, d, d])[..., None], 1.0)
for rm in (1, 2, 3, 4, 6, 8, 12, 16, 32):
    pb.set_schedule(rm)
    for _ in range(3):
        pb.simulate_logp(P, Y, lp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pb.simulate_logp(P, Y, lp)
    e1.record()
    torch.cuda.synchronize()
    print(f"refill_min {rm:2d}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
# sorted inputs (by k_abs + CL): how much is left of the divergence cost?
order = np.argsort(th[0] + th[1])
Ps = torch.from_numpy(np.ascontiguousarray(th[:, order])).cuda()
pb.set_schedule(3)
for name, PP in (("unsorted", P), ("sorted by k_abs+CL", Ps)):
    for _ in range(3):
        pb.simulate_logp(PP, Y, lp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pb.simulate_logp(PP, Y, lp)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
