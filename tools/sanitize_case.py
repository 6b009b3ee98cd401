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
        assert nf == 0 and nf2 == 0 and nf3 == 0 and np.all(np.isfinite(lp)) and np.array_equal(lp, lp2)
    pb.close()
print("sanitize case ok")
