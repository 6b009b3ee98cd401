#!/usr/bin/env python
"""Time the host-buffer call of the C2 workload (pinned P -> Y, logp on the host): trajectories/s and effective D2H GB/s.

Tuning aid:  PV_HOST_CHUNK=<rows> python tools/time_e2e.py
"""
import importlib
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rt = importlib.import_module("parameter-variability_b200.runtime")
B, T = 1_000_000, 50
np.random.seed(1234)
th = np.stack([np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B) for _ in range(2)])
pb = rt.Problem("simple_pk")
pb.set_solver(method=rt.METHOD_AUTO, rtol=1e-7, atol=1e-10)
pb.set_columns(["k_abs", "CL"])
pb.set_timepoints(np.linspace(0, 100, T))
pb.set_observables(["y_gut", "y_cent", "y_peri"])
d = np.ones((T, 3))
pb.set_data(np.stack([d, d, d])[..., None], 1.0)
P = rt.pinned_empty((2, B))
P[:] = th
Y = rt.pinned_empty((B, T, 3))
lp = rt.pinned_empty((B,))
for _ in range(2):
    pb.simulate_host(P, B, Y=Y, logp=lp)
n = 6
t0 = time.perf_counter()
for _ in range(n):
    pb.simulate_host(P, B, Y=Y, logp=lp)
dt = (time.perf_counter() - t0) / n
print(f"chunk={os.environ.get('PV_HOST_CHUNK', 'default')}: {dt * 1e3:.2f} ms per pass, {B / dt / 1e6:.2f} M trajectories/s, "
      f"{(B * (T * 3 + 1) * 8) / dt / 1e9:.1f} GB/s D2H, logp_sum={float(lp.sum()):.4f}", flush=True)
