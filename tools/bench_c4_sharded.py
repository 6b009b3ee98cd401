#!/usr/bin/env python
"""BASELINE config C4 measurement: icg_body_flat (stiff, RODAS4 + forward sensitivities), hierarchical objective,
individuals sharded over the ranks, ONE NCCL all-reduce of [n_chains, 1 + 2P] per evaluation.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      tools/bench_c4_sharded.py --individuals-per-gpu 2000 --chains 8 --evals 20

Weak scaling (individuals per GPU fixed).  Rank 0 prints one JSON line; timing = CUDA-synchronised wall time of the
whole evaluation (kernel + D2H + host population terms + all-reduce), max over ranks.
"""
import argparse
import importlib
import json
import os
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
    ap.add_argument("--individuals-per-gpu", type=int, default=2000)
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--evals", type=int, default=20)
    ap.add_argument("--check", action="store_true", help="rank 0 also evaluates the unsharded objective and compares")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import parvar_oracle as oracle
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = importlib.import_module(PKG + ".runtime")
    po = importlib.import_module(PKG + ".optimization.petab_optimization")
    dd = importlib.import_module(PKG + ".distributed")

    C, n_loc = args.chains, args.individuals_per_gpu
    n_ind = n_loc * world
    T = 20
    t = np.linspace(0, 100, T)
    obs = ["Cre_plasma_icg", "Cgi_plasma_icg"]
    idx = [oracle.ICG_STATES.index(o) for o in obs]
    truth = oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0}, t)[:, idx]
    first, stop = dd.shard_range(n_ind, rank, world)
    rs = np.random.RandomState(1234)                              # every rank draws the full set, keeps its slice
    y_all = truth[None] * (1 + 0.05 * rs.normal(size=(n_ind,) + truth.shape))
    m1, s1 = oracle.lognorm_par_transform(75.0, 10.0)
    m2, s2 = oracle.lognorm_par_transform(0.0369598840327503, 0.01)
    theta_all = np.stack([rs.lognormal(m1, s1, size=(C, n_ind)), rs.lognormal(m2, s2, size=(C, n_ind))], axis=2)   # [C, n_ind, 2]
    mu = np.tile(np.array([m1, m2]), (C, 1))
    omega = np.tile(np.array([s1, s2]), (C, 1))

    def make(first_, stop_):
        pb = rt.Problem("icg_body_flat", device=local)
        pb.set_solver(rtol=1e-7, atol=1e-11)
        pb.set_value("IVDOSE_icg", 10.0)
        pb.set_columns(["BW", "LI__ICGIM_Vmax"])
        pb.set_timepoints(t)
        pb.set_observables(obs)
        y = y_all[first_:stop_]
        pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(len(y))], axis=-1), 1e-4)
        return po.HierarchicalObjective(pb, n_ind, mu_prior=[(m1, 0.5), (m2, 0.5)], omega_prior=[0.5, 0.5],
                                        first_individual=first_, n_local=stop_ - first_)

    H = make(first, stop)
    theta = np.ascontiguousarray(theta_all[:, first:stop])

    def local_eval(_theta_unused, first_, n_local_):
        lp, dth, dmu, dom = H.local_terms(theta, mu, omega)
        # pack: ShardedLikelihood sums logp over individuals itself, so hand it per-chain values spread over one column
        return lp, dth, dmu, dom

    def evaluate():
        lp, dth, dmu, dom = H.local_terms(theta, mu, omega)
        buf = np.ascontiguousarray(np.concatenate([lp[:, None], dmu, dom], axis=1))          # [C, 1 + 2P]
        if world > 1:
            tb = torch.from_numpy(buf).cuda()
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            buf = tb.cpu().numpy()
        logp, gmu, gom = H.finish(buf[:, 0], buf[:, 1:3], buf[:, 3:5], mu, omega)
        return logp, dth, gmu, gom

    for _ in range(3):
        out = evaluate()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.evals):
        out = evaluate()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    line = {"config": "C4 icg_body_flat hierarchical logp+grad, individuals sharded, 1 all-reduce per evaluation",
            "n_gpus": world, "chains": C, "individuals_per_gpu": n_loc, "individuals": n_ind,
            "evals_per_s": args.evals / dt, "chain_evals_per_s": C * args.evals / dt, "ms_per_eval": 1e3 * dt / args.evals,
            "trajectories_per_s": C * n_ind * args.evals / dt, "kernel_ms_last": H.problem.last_kernel_ms,
            "logp_chain0": float(out[0][0])}
    if args.check and rank == 0 and world > 1:
        Hfull = make(0, n_ind)
        lp, dth, dmu, dom = Hfull(theta_all, mu, omega)
        line["unsharded_logp_chain0"] = float(lp[0])
        line["sharded_matches_unsharded"] = bool(np.allclose(lp, out[0], rtol=1e-12) and np.allclose(dmu, out[2], rtol=1e-9))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
