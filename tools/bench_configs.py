#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): BASELINE configs C3 / C4 and the per-kernel throughputs.
Prints one JSON line per measurement; results are copied into profiles/ per round.

  python tools/bench_configs.py [--reps 20] [--only c3,c4,icg_fwd,pk_grad]
"""
import argparse
import importlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
PKG = "parameter-variability_b200"
MU, SG = -0.3465735902799726, 0.8325546111576977


def time_launches(fn, reps):
    import torch
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
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--layout", type=int, default=0, help="pv_problem_set_layout for the Rosenbrock sensitivity kernels: 0 auto, 1 lane, 2 split")
    ap.add_argument("--locality", type=int, default=0, help="pv_problem_set_locality: 0 auto, 1 off, 2 on")
    ap.add_argument("--lockstep", type=int, default=0, help="pv_problem_set_lockstep: 0 auto, 1 off, 2 on")
    ap.add_argument("--sizes", default="1000,16000", help="individuals per chain for the icg cases")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    import torch
    import parvar_oracle as oracle
    rt = importlib.import_module(PKG + ".runtime")
    opt = importlib.import_module(PKG + ".optimization.petab_optimization")
    bc = importlib.import_module(PKG + ".bench_cases")
    peak64 = rt.fma_peak_tflops(0, True)
    stream = torch.cuda.current_stream().cuda_stream

    def emit(name, ms, B, pb, steps, step_key, dense_key, T, extra=None):
        att = int(steps.sum())
        flops = att * pb.flops(step_key) + B * T * pb.flops(dense_key)
        line = {"config": name, "ms_per_eval": ms, "trajectories": B, "trajectories_per_s": B / (ms * 1e-3),
                "attempted_steps_per_trajectory": att / B, "max_steps": int(steps.sum(dim=0).max()),
                "flops_per_step": pb.flops(step_key), "achieved_tflops": flops / (ms * 1e-3) / 1e12,
                "fp64_peak_tflops": peak64, "frac_of_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak64}
        line.update(extra or {})
        line.update({"locality": args.locality, "lockstep": args.lockstep})
        print(json.dumps(line), flush=True)

    # ---- C3: simple_pk, 8 chains x 1000 individuals, logp + grad (forward sensitivities), T = 20 ----------------
    if not only or "c3" in only:
        n_chains, n_ind, T = 8, 1000, 20
        t = np.linspace(0, 100, T)
        rs = np.random.RandomState(1234)
        th_true = rs.lognormal(MU, SG, size=(2, n_ind))
        y = np.stack([oracle.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in zip(*th_true)])
        y = y * (1 + 0.05 * rs.normal(size=y.shape))
        stats = np.stack([opt.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1)
        pb = rt.Problem("simple_pk")
        pb.set_solver(rtol=1e-8, atol=1e-11)
        pb.set_columns(["k_abs", "CL"])
        pb.set_timepoints(t)
        pb.set_observables(["y_gut", "y_cent", "y_peri"])
        pb.set_data(stats, 0.05)
        pb.set_locality(args.locality)
        pb.set_lockstep(args.lockstep)
        B = n_chains * n_ind
        th = np.concatenate([np.random.RandomState(1234 + c).lognormal(MU, SG, size=(2, n_ind)) for c in range(n_chains)], axis=1)
        P = torch.from_numpy(th).cuda()
        lp = torch.empty(B, dtype=torch.float64, device="cuda")
        g = torch.empty((2, B), dtype=torch.float64, device="cuda")
        seg = torch.empty(n_chains, dtype=torch.float64, device="cuda")
        steps = torch.zeros((2, B), dtype=torch.int32, device="cuda")
        ms = time_launches(lambda: pb.logp_grad(P, lp, g, seg_len=n_ind, logp_seg=seg, steps=steps, stream=stream), args.reps)
        emit("C3 simple_pk logp+grad 8 chains x 1000 individuals (T=20, rtol 1e-8)", ms, B, pb, steps, "step_dopri5_sens",
             "dense_dopri5_sens", T, {"logp_grad_evals_per_s": n_chains / (ms * 1e-3)})
        for mult in (16, 128):
            Bm = B * mult
            Pm = P.repeat(1, mult)
            lpm = torch.empty(Bm, dtype=torch.float64, device="cuda")
            gm = torch.empty((2, Bm), dtype=torch.float64, device="cuda")
            stm = torch.zeros((2, Bm), dtype=torch.int32, device="cuda")
            ms = time_launches(lambda: pb.logp_grad(Pm, lpm, gm, steps=stm, stream=stream), max(args.reps // 4, 3))
            emit(f"C3 x{mult}: simple_pk logp+grad, {Bm} trajectories", ms, Bm, pb, stm, "step_dopri5_sens", "dense_dopri5_sens", T)

    # ---- icg forward + C4-style logp+grad -------------------------------------------------------------------------
    if not only or "icg" in only or "c4" in only:
        T = 20
        t = np.linspace(0, 100, T)
        obs = ["Cre_plasma_icg", "Cgi_plasma_icg"]
        idx = [oracle.ICG_STATES.index(o) for o in obs]
        truth = oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0}, t)[:, idx]
        rs = np.random.RandomState(1234)
        m1, s1 = oracle.lognorm_par_transform(75.0, 10.0)
        m2, s2 = oracle.lognorm_par_transform(0.0369598840327503, 0.01)
        for n_chains, n_ind in [(8, int(x)) for x in args.sizes.split(",")]:
            B = n_chains * n_ind
            th = np.stack([rs.lognormal(m1, s1, B), rs.lognormal(m2, s2, B)])
            y = truth[None] * (1 + 0.05 * rs.normal(size=(n_ind,) + truth.shape))
            stats = np.stack([opt.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1)
            pb = rt.Problem("icg_body_flat")
            pb.set_solver(rtol=bc.C4.rtol, atol=bc.C4.atol)
            pb.set_locality(args.locality)
            pb.set_lockstep(args.lockstep)
            pb.set_value("IVDOSE_icg", 10.0)
            pb.set_columns(["BW", "LI__ICGIM_Vmax"])
            pb.set_timepoints(t)
            pb.set_observables(obs)
            pb.set_data(stats, 1e-4)
            pb.set_layout(args.layout)
            P = torch.from_numpy(th).cuda()
            Y = torch.empty((B, T, 2), dtype=torch.float64, device="cuda")
            lp = torch.empty(B, dtype=torch.float64, device="cuda")
            g = torch.empty((2, B), dtype=torch.float64, device="cuda")
            seg = torch.empty(n_chains, dtype=torch.float64, device="cuda")
            steps = torch.zeros((2, B), dtype=torch.int32, device="cuda")
            ms = time_launches(lambda: pb.simulate(P, Y, steps=steps, stream=stream), max(args.reps // 4, 3))
            emit(f"icg_body_flat forward RODAS4, {B} trajectories (T=20, rtol {bc.C4.rtol:g})", ms, B, pb, steps, "step_rodas4", "dense_rodas4", T)
            ms = time_launches(lambda: pb.logp(P, lp, steps=steps, stream=stream), max(args.reps // 4, 3))
            emit(f"icg_body_flat logp RODAS4, {B} trajectories", ms, B, pb, steps, "step_rodas4", "dense_rodas4", T)
            ms = time_launches(lambda: pb.logp_grad(P, lp, g, seg_len=n_ind, logp_seg=seg, steps=steps, stream=stream),
                               max(args.reps // 4, 3))
            emit(f"C4 icg_body_flat logp+grad {n_chains} chains x {n_ind} individuals", ms, B, pb, steps, "step_rodas4_sens",
                 "dense_rodas4_sens", T, {"logp_grad_evals_per_s": n_chains / (ms * 1e-3)})


if __name__ == "__main__":
    main()
