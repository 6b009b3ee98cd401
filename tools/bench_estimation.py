#!/usr/bin/env python
"""Estimation throughput: Q simple_pk PEtab problems (the reference's sweep cell: 2 groups, 20 samples, 20 timepoints,
cv 0.05, exact priors) optimised (multistart) and sampled (adaptive Metropolis + parallel tempering, the reference's
settings petab_optimization.py:294-306) TOGETHER.  Prints one JSON line."""
import argparse
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
PKG = "parameter-variability_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=64)
    ap.add_argument("--starts", type=int, default=100)
    ap.add_argument("--samples", type=int, default=5000)
    ap.add_argument("--chains", type=int, default=4)
    args = ap.parse_args()
    import parvar_oracle as oracle
    po = importlib.import_module(PKG + ".optimization.petab_optimization")
    drv = importlib.import_module(PKG + ".optimization.driver")
    t = np.linspace(0, 100, 21)
    obs = ["y_cent", "y_gut", "y_peri"]
    cols = [oracle.SIMPLE_PK.states.index(o) for o in obs]
    truths = {"MALE": (0.5, 1.0), "FEMALE": (1.0, 0.5)}           # (k_abs, CL) locations of factories/simple_pk.py:19-44
    priors = {"k_abs_MALE": (0.5, 1.0), "k_abs_FEMALE": (1.0, 1.0), "CL_MALE": (1.0, 1.0), "CL_FEMALE": (0.5, 1.0)}
    problems = []
    for q in range(args.problems):
        rs = np.random.RandomState(1234 + q)
        groups = []
        for gid, (ka, cl) in truths.items():
            th = np.stack([rs.lognormal(np.log(ka), 0.3, 20), rs.lognormal(np.log(cl), 0.3, 20)])
            y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t)[:, cols] for a, c in zip(*th)])
            groups.append(po.GroupData(gid, y * (1 + 0.05 * rs.normal(size=y.shape))))
        problems.append(groups)
    batch = po.PetabObjectiveBatch("simple_pk", ["k_abs", "CL"], problems, t, obs, sigma=1.0, priors=[priors] * args.problems)
    d = drv.GpuPetabSampler(batch)                                 # start points uniform in [0, 1e8] like the reference
    t0 = time.perf_counter()
    res = d.run_optimization(n_starts=args.starts, maxiter=200)
    t_opt = time.perf_counter() - t0
    n_opt = batch.n_evals
    t0 = time.perf_counter()
    sr = d.bayesian_sampler(n_samples=args.samples, n_chains=args.chains)
    t_smp = time.perf_counter() - t0
    n_smp = batch.n_evals - n_opt
    print(json.dumps({"config": f"{args.problems} simple_pk PEtab problems x ({args.starts} starts, {args.chains} x {args.samples} samples)",
                      "optim_s": t_opt, "sampler_s": t_smp, "problems_per_hour": 3600 * args.problems / (t_opt + t_smp),
                      "objective_evals_optim": n_opt, "objective_evals_sampler": n_smp,
                      "objective_evals_per_s": (n_opt + n_smp) / (t_opt + t_smp),
                      "best_fval_median": float(np.median(res.fval[:, 0])), "converged_best": int(res.converged[:, 0].sum()),
                      "ess_median": float(np.median(sr.effective_sample_size)),
                      "cold_acceptance_median": float(np.median(sr.acceptance_rate[:, 0]))}), flush=True)


if __name__ == "__main__":
    main()
