#!/usr/bin/env python
"""Which integrator tolerance keeps logp / gradient of the bench's C3 / C4 cases inside the 1e-6 bar?

Runs the DEVICE integrator templates compiled for the host (tests/host/harness.py: the same pv_dopri5 / pv_rosenbrock
code and generated model functions the kernels instantiate) on the seeded draws of ``bench_cases`` and compares with the
oracle at rtol 1e-12.  CPU only; the GPU tests (tests/test_gpu_parity.py::test_bench_logp_grad_settings_meet_the_bar)
assert the chosen setting on the box.

    python tools/tolerance_scan.py C4 --pairs 128
"""
from __future__ import annotations

import argparse
import importlib
import sys
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests" / "host"):
    sys.path.insert(0, str(p))
import harness            # noqa: E402
import parvar_oracle as oracle   # noqa: E402

bc = importlib.import_module("parameter-variability_b200.bench_cases")


def _oracle_one(args):
    case_name, over, y = args
    case = {c.name.split()[0]: c for c in bc.CASES}[case_name]
    m = oracle.MODELS[case.model]
    idx = [m.states.index(o) for o in case.observables]
    Yo, So = oracle.simulate_sens(case.model, over, case.tgrid, list(case.columns))
    f, s = Yo[:, idx], So[:, idx, :]
    mass = np.abs(((y - f) / case.sigma ** 2)[..., None] * s).sum(axis=(0, 1))
    return oracle.loglik_normal(f, y, case.sigma), oracle.loglik_normal_grad(f, s, y, case.sigma), mass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["C3", "C4"])
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--tols", default="1e-7:1e-11,3e-8:1e-11,1e-8:1e-11,1e-8:1e-12")
    a = ap.parse_args()
    case = {c.name.split()[0]: c for c in bc.CASES}[a.case]
    t = case.tgrid
    m = oracle.MODELS[case.model]
    idx = [m.states.index(o) for o in case.observables]
    truth = oracle.simulate(case.model, {**case.fixed, **case.mean_parameters()}, t)[:, idx]
    y, theta = case.inputs(truth, bc.N_INDIVIDUALS_PER_GPU)
    n_i = a.pairs // bc.N_CHAINS
    pairs = [(c, i) for c in range(bc.N_CHAINS) for i in range(n_i)]
    overs = [{**case.fixed, **{k: float(theta[c, i, p]) for p, k in enumerate(case.columns)}} for c, i in pairs]
    with Pool() as pool:
        ref = pool.map(_oracle_one, [(a.case, o, y[i]) for o, (c, i) in zip(overs, pairs)])
    meta = harness.meta(case.model)
    hidx = [meta["states"].index(o) for o in case.observables]
    nx, method = meta["NX"], (1 if case.method == "DOPRI5" else 2)
    for tol in a.tols.split(","):
        rtol, atol = (float(v) for v in tol.split(":"))
        w_lp = w_c = w_m = w_wc = 0.0
        cancel = []
        steps = []
        for (c, i), over, (rl, rg, mass) in zip(pairs, overs, ref):
            rc, Z, (acc, rej) = harness.simulate(case.model, over, t, rtol, atol, method=method, sens=True)
            assert rc == 0
            steps.append(acc + rej)
            f = Z[:, :nx][:, hidx]
            s = np.stack([Z[:, nx * (1 + k):nx * (2 + k)][:, hidx] for k in range(len(case.columns))], axis=-1)
            lp, g = oracle.loglik_normal(f, y[i], case.sigma), oracle.loglik_normal_grad(f, s, y[i], case.sigma)
            w_lp = max(w_lp, abs(lp - rl) / abs(rl))
            w_c = max(w_c, float(np.max(np.abs(g - rg) / np.abs(rg))))
            w_m = max(w_m, float(np.max(np.abs(g - rg) / mass)))
            well = mass <= 10.0 * np.abs(rg)           # components that do not cancel by more than 10x
            cancel.append(float(np.max(mass / np.abs(rg))))
            if well.any():
                w_wc = max(w_wc, float(np.max((np.abs(g - rg) / np.abs(rg))[well])))
        print(f"{case.name}: rtol {rtol:g} atol {atol:g}: attempted steps mean {np.mean(steps):.0f} max {max(steps)}; "
              f"logp rel {w_lp:.2e}; gradient component-wise rel {w_c:.2e} (components cancelling < 10x: {w_wc:.2e}; worst "
              f"cancellation {max(cancel):.0f}x); gradient / sum|terms| {w_m:.2e}", flush=True)


if __name__ == "__main__":
    main()
