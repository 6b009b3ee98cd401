#!/usr/bin/env python
"""bench.py's logp + gradient leg (bench_cases C3 / C4, 8 chains x N individuals) taken apart on one GPU: the
sensitivity kernel alone (with its step counts) against the whole hierarchical evaluation (kernel + epilogue + hyper-prior
kernels).  Tuning aid:  PARVAR_B200_LIB=build/variants/<name>.so python tools/time_hier.py [--cases C3] [--n 1000]
"""
import argparse
import importlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
PKG = "parameter-variability_b200"
import torch  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="C3,C4")
    ap.add_argument("--n", type=int, default=1000, help="individuals per chain")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--group-lanes", type=int, default=0)
    ap.add_argument("--layout", type=int, default=0)
    args = ap.parse_args()
    rt = importlib.import_module(PKG + ".runtime")
    po = importlib.import_module(PKG + ".optimization.petab_optimization")
    bc = importlib.import_module(PKG + ".bench_cases")
    C = bc.N_CHAINS
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    for case in bc.CASES:
        if case.name.split()[0] not in args.cases.split(","):
            continue
        P = len(case.columns)
        t = case.tgrid
        pb = rt.Problem(case.model)
        pb.set_solver(rtol=case.rtol, atol=case.atol)
        for k, v in case.fixed.items():
            pb.set_value(k, v)
        pb.set_columns(list(case.columns))
        pb.set_timepoints(t)
        pb.set_observables(list(case.observables))
        pb.set_group_lanes(args.group_lanes)
        pb.set_layout(args.layout)
        mean = torch.tensor([[m] for m, _ in case.population], dtype=torch.float64, device=dev)
        Y1 = torch.empty((1, len(t), len(case.observables)), dtype=torch.float64, device=dev)
        pb.simulate(mean, Y1)
        torch.cuda.synchronize()
        truth = Y1[0].cpu().numpy()
        pop = case.population_ln()
        mu = torch.tensor([[m for m, _ in pop]] * C, dtype=torch.float64, device=dev)
        omega = torch.tensor([[s for _, s in pop]] * C, dtype=torch.float64, device=dev)
        n = args.n
        y, theta = case.inputs(truth, n)
        pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n)], axis=-1), case.sigma)
        th = torch.from_numpy(np.ascontiguousarray(theta.transpose(2, 0, 1))).to(dev)     # [P, C, n]
        H = po.HierarchicalObjectiveDevice(pb, C, n, [(m, 0.5) for m, _ in pop], [0.5] * P)
        ms_h = timed(lambda: H(th, mu, omega), args.reps)
        B = C * n
        lp = torch.empty(B, dtype=torch.float64, device=dev)
        g = torch.empty((P, B), dtype=torch.float64, device=dev)
        steps = torch.zeros((2, B), dtype=torch.int32, device=dev)
        Pm = th.reshape(P, B)
        ms_k = timed(lambda: pb.logp_grad(Pm, lp, g, steps=steps, stream=stream), args.reps)
        tot = steps.sum(dim=0)
        print(json.dumps({"case": case.name, "lib": str(rt.LIB_PATH.name), "individuals": n, "ms_hier_eval": ms_h, "ms_sens_kernel": ms_k,
                          "max_steps": int(tot.max()), "mean_steps": float(tot.double().mean()),
                          "us_per_step_of_longest": 1e3 * ms_k / int(tot.max()), "logp_sum": float(lp.sum()),
                          "grad_sum": [float(x) for x in g.sum(dim=1)]}), flush=True)
        pb.close()


if __name__ == "__main__":
    main()
