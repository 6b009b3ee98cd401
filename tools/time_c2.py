#!/usr/bin/env python
"""Time the C2 kernel (simple_pk, 1e6 draws x 50 timepoints, trajectories + logp) on the device: ms per launch.

Tuning aid for build variants:  PARVAR_B200_LIB=build/variants/<name>.so python tools/time_c2.py [--refill N]
"""
import argparse
import importlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--refill", type=int, default=0)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--locality", type=int, default=0, help="0 auto, 1 off, 2 on (pv_problem_set_locality)")
ap.add_argument("--lockstep", type=int, default=0, help="0 auto, 1 off, 2 on (pv_problem_set_lockstep)")
ap.add_argument("--mode", default="traj_logp", choices=["traj_logp", "logp", "traj"])
args = ap.parse_args()
rt = importlib.import_module("parameter-variability_b200.runtime")
B, T = 1_000_000, 50
np.random.seed(1234)
th = np.stack([np.random.lognormal(-0.3465735902799726, 0.8325546111576977, B) for _ in range(2)])
P = torch.from_numpy(th).cuda()
Y = torch.empty((B, T, 3), dtype=torch.float64, device="cuda")
lp = torch.empty(B, dtype=torch.float64, device="cuda")
pb = rt.Problem("simple_pk")
pb.set_solver(method=rt.METHOD_DOPRI5, rtol=1e-7, atol=1e-10)
pb.set_columns(["k_abs", "CL"])
pb.set_timepoints(np.linspace(0, 100, T))
pb.set_observables(["y_gut", "y_cent", "y_peri"])
d = np.ones((T, 3))
pb.set_data(np.stack([d, d, d])[..., None], 1.0)
if args.refill:
    pb.set_schedule(args.refill)
pb.set_locality(args.locality)
pb.set_lockstep(args.lockstep)
st = torch.zeros((2, B), dtype=torch.int32, device="cuda")
status = torch.zeros(B, dtype=torch.int32, device="cuda")
run = {"traj_logp": lambda: pb.simulate_logp(P, Y, lp, status=status, steps=st), "logp": lambda: pb.logp(P, lp), "traj": lambda: pb.simulate(P, Y)}[args.mode]
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{os.environ.get('PARVAR_B200_LIB') or 'default'} refill={args.refill or 'default'} locality={args.locality} lockstep={args.lockstep} mode={args.mode}: "
      f"{e0.elapsed_time(e1) / args.reps:.3f} ms  logp_sum={float(lp.sum()):.6f} steps/traj={float(st.sum()) / B:.1f} failed={int((status != 0).sum())} "
      f"Y_bits_checksum={int(Y.view(torch.int64).sum()) if args.mode != 'logp' else 0}", flush=True)
