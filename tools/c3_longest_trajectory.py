#!/usr/bin/env python
"""Longest trajectory (attempted DOPRI5 + sensitivity steps, per-lane stepper compiled for the host) among bench.py's C3 draws at
1000 and at 8000 individuals per chain: a one-wave evaluation lasts as long as its longest trajectory, so the weak-scaling time of
the C3 leg at 8 ranks (8 x 1000 individuals per chain) follows the maximum over 64 000 draws, not a collective.  CPU only."""
import sys, importlib, numpy as np, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from host import harness as hh
bc = importlib.import_module('parameter-variability_b200.bench_cases')
case = bc.C3
t = case.tgrid
truth = np.zeros((len(t), 3))
out = {}
for n_ind in (1000, 8000):
    y, theta = case.inputs(truth, n_ind)      # theta [C, n_ind, P]
    th = theta.reshape(-1, 2)
    # cheap screen: per-lane DOPRI5 with sensitivities; only the stiffest ~2% of the draws can be the maximum
    key = th.max(axis=1)
    idx = np.argsort(key)[-max(200, len(th)//50):]
    t0 = time.time()
    steps = []
    for k in idx:
        rc, Y, (a, r) = hh.simulate('simple_pk', {'k_abs': th[k,0], 'CL': th[k,1]}, t, case.rtol, case.atol, method=1, sens=True)
        steps.append(a + r)
    out[n_ind] = (int(np.max(steps)), len(idx), round(time.time()-t0,1))
    print(n_ind, 'individuals per chain:', len(th), 'trajectories, max attempted steps (per-lane DOPRI5 + sens) =', out[n_ind])
