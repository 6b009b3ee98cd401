#!/usr/bin/env python
"""bench.py - headline benchmark of the parvar hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): simple_pk, 1 000 000 log-normal
parameter draws (k_abs, CL; mean = sd = 1, np.random.seed(1234 + rank)) x 50 timepoints on [0, 100],
3 observables.  One step = one pass over the batch: every trajectory is integrated once, its 50 x 3
concentrations are written to HBM (Y[B][50][3], 1.2 GB) and its Gaussian log-likelihood against one
fixed synthetic data vector is reduced in the same kernel (logp[B]).

  value      trajectories/s, whole job (all ranks), inputs resident in HBM, CUDA events, max over ranks
  e2e        same metric through the C ABI's host-buffer call (pinned host P -> Y, logp on the host)
  roofline   fp64 FMA pipe: algorithmic flops (code-generator flop counts x steps counted by the kernel)
             / kernel time, against the fp64 FMA peak measured in this run by pv_fma_peak_tflops;
             plus the HBM view (algorithmic bytes / time against MEASURED_PEAKS.json hbm_gbs)
  cpu_baseline  the oracle's C restatement (oracle/, OpenMP over all host cores) on a bounded sample

`--impl reference` times the CPU arm only (the reference's libroadrunner/AMICI cannot be installed here:
DESIGN.md section "Oracle"; the arm is the oracle port, kind = "port").
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "parameter-variability_b200"

B_PER_GPU = 1_000_000
T_POINTS = 50
MU_LN, SIGMA_LN = -0.3465735902799726, 0.8325546111576977   # lognorm_par_transform(1, 1), petab_factory.py:94-98
RTOL, ATOL = 1e-7, 1e-10     # integrator tolerances at which the parity tests pass (rtol 1e-6 / atol 1e-9 vs oracle)
OBS = ["y_gut", "y_cent", "y_peri"]


def draws(rank: int, B: int) -> np.ndarray:
    np.random.seed(1234 + rank)
    k_abs = np.random.lognormal(MU_LN, SIGMA_LN, B)
    CL = np.random.lognormal(MU_LN, SIGMA_LN, B)
    return np.stack([k_abs, CL])


def c1_data(truth: np.ndarray) -> np.ndarray:
    """C1's synthetic data vector: default-parameter trajectory x (1 + 0.05 eps), eps ~ seed 1234.
    `truth` [T, 3] comes from the arm that is being timed (GPU arm: one trajectory through the CUDA path; CPU arm: the
    oracle's closed form), so the GPU arm never touches oracle/."""
    rs = np.random.RandomState(1234)
    return truth * (1.0 + 0.05 * rs.normal(size=truth.shape))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
def cpu_baseline(sample: int, threads: int | None = None) -> dict:
    """Oracle port timed on the host cores: the C restatement (oracle/parvar_oracle.c, variable-order BDF like
    CVODE, the reference's tolerances rtol = atol = 1e-6, petab_factory.py:183-189) with OpenMP over all host
    threads.  Same work per trajectory as one step of the GPU arm: 50 x 3 outputs + the Gaussian logp."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_c
    tgrid = np.linspace(0.0, 100.0, T_POINTS)
    th = draws(0, sample)
    cores = threads or os.cpu_count() or 1
    import parvar_oracle as oracle
    data = c1_data(oracle.simple_pk_exact({}, tgrid))
    oracle_c.load()
    t0 = time.perf_counter()
    _, lp, nf, st = oracle_c.simulate_batch("simple_pk", {"k_abs": th[0], "CL": th[1]}, tgrid, OBS, rtol=1e-6, atol=1e-6,
                                            threads=cores, data=data, sigma=1.0)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "sample": f"first {sample} of the {B_PER_GPU} draws; C restatement oracle/parvar_oracle.c (variable-order BDF + "
                      f"Newton + dense LU as CVODE, rtol=atol=1e-6 as petab_factory.py:183-189), OpenMP {cores} threads; "
                      f"{st[0] / sample:.0f} steps/trajectory",
            "seconds": dt, "failed": nf}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(max(sample // 10, 100))
    for _ in range(args.steps):
        vals.append(cpu_baseline(sample))
    best = vals[-1]
    v = float(np.mean([x["value"] for x in vals]))
    ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
    line = {"impl": "reference", "metric": "ODE trajectories/s (simple_pk forward simulation + logp)", "value": v,
            "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {k: best[k] for k in ("unit", "cores", "kind", "sample")} | {"value": v},
            "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "C2: simple_pk batched forward simulation + Gaussian logp, "
                        f"{B_PER_GPU} log-normal (k_abs, CL) draws x {T_POINTS} timepoints x 3 observables per GPU",
            "trajectories_per_gpu": B_PER_GPU, "timepoints": T_POINTS, "observables": OBS,
            "integrator": "METHOD_AUTO: DOPRI5 adaptive with dense output, stiffness detection; stiff rows are re-run with RODAS4 (none in this cloud)",
            "schedule": "device-side locality sort of the draws (stable radix sort, inside the timed region) + lockstep warps (one step size per warp)", "rtol": RTOL, "atol": ATOL,
            "l2": "outputs (1.2 GB per step) exceed the 126 MB L2, so every step streams to HBM",
            "sharding": f"independent draws, {n_gpus} rank(s), no data-path collective"}


# ----------------------------------------------------------------------------------------------------------
def _timed(torch, dist, world, fn, reps):
    """reps calls of fn between barrier + synchronize, CUDA events on the current stream, max over ranks -> ms per call."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sampler_leg(fn) -> dict:
    try:
        return fn()
    except Exception as exc:      # noqa: BLE001 - secondary measurement
        return {"error": f"{type(exc).__name__}: {exc}"}


def _hmc_leg(H, x0, world, dist) -> dict:
    """optimization/hmc.py::hmc_device: drift kernel -> evaluation -> kick kernel per leapfrog step, nothing on the host."""
    hm = importlib.import_module(PKG + ".optimization.hmc")
    hm.hmc_device(H, x0, n_samples=2, n_warmup=2, n_leapfrog=4, step_size=1e-3, seed=1, keep_theta=False)
    if world > 1:
        dist.barrier()
    t_h = time.perf_counter()
    res = hm.hmc_device(H, x0, n_samples=10, n_warmup=10, n_leapfrog=10, step_size=1e-3, seed=1, keep_theta=False)
    dt_h = time.perf_counter() - t_h
    return {"leapfrog_steps": int(res.n_grad_evals), "ms_per_leapfrog_step": 1e3 * dt_h / res.n_grad_evals,
            "leapfrog_steps_per_s": res.n_grad_evals / dt_h, "mean_accept": float(np.mean(res.accept_rate)),
            "what": "20 HMC iterations x ~10 leapfrog steps of all 8 chains, sampler state device resident, wall clock of the whole "
                    "call (random-number upload and result download included)"}


def _nuts_leg(H, x0) -> dict:
    """optimization/nuts.py::nuts_device: the sampler BASELINE config 3 names, tree on the GPU; the host reads 8 flags per
    leapfrog step, so the number includes one stream synchronisation per step."""
    nt = importlib.import_module(PKG + ".optimization.nuts")
    nt.nuts_device(H, x0, n_samples=1, n_warmup=2, step_size=1e-3, max_depth=3, seed=1)
    t_n = time.perf_counter()
    rn = nt.nuts_device(H, x0, n_samples=6, n_warmup=6, step_size=1e-3, max_depth=4, seed=1)
    dt_n = time.perf_counter() - t_n
    return {"leapfrog_steps": int(rn.n_grad_evals), "ms_per_leapfrog_step": 1e3 * dt_n / rn.n_grad_evals,
            "leapfrog_steps_per_s": rn.n_grad_evals / dt_n, "mean_tree_depth": float(np.mean(rn.tree_depth)),
            "what": "12 NUTS iterations (max depth 4) of all 8 chains, tree state device resident, wall clock of the whole call"}


def logp_grad_leg(rt, torch, dist, rank: int, world: int, local: int) -> dict:
    """Second half of BASELINE.json's metric: logp + gradient evaluations/s (configs[2] "C3" and [3] "C4"), at exactly the
    settings of ``bench_cases`` (the ones tests/test_gpu_parity.py::test_bench_logp_grad_settings_meet_the_bar verifies).

    Hierarchical objective over 8 chains x n individuals, individuals sharded over the ranks, device resident
    (``HierarchicalObjectiveDevice``): sensitivity kernel -> epilogue kernel writing [8, 5] -> ONE NCCL all-reduce of that
    buffer on the same stream -> hyper-prior kernel; theta, gradients and results never leave the GPU.
      weak             1000 individuals per GPU (the config BASELINE names, per GPU)
      strong_8x1000    the config BASELINE literally names, 8 x 1000 individuals in total, split over the ranks
      strong_8x16000   a size that fills the GPUs: 8 x 16 000 individuals in total, split over the ranks
      e2e              weak size through host buffers: pinned theta -> H2D, evaluation, D2H of [8, 5] and the gradients
    Data are synthetic: the trajectory at the population means (through the CUDA path) x (1 + 0.05 eps) per individual.
    """
    po = importlib.import_module(PKG + ".optimization.petab_optimization")
    dd = importlib.import_module(PKG + ".distributed")
    bc = importlib.import_module(PKG + ".bench_cases")
    C = bc.N_CHAINS
    dev = torch.device("cuda", local)
    out = {}
    for case in bc.CASES:
        P = len(case.columns)
        t = case.tgrid
        pb = rt.Problem(case.model, device=local)
        pb.set_solver(rtol=case.rtol, atol=case.atol)
        for k, v in case.fixed.items():
            pb.set_value(k, v)
        pb.set_columns(list(case.columns))
        pb.set_timepoints(t)
        pb.set_observables(list(case.observables))
        mean = torch.tensor([[m] for m, _ in case.population], dtype=torch.float64, device=dev)
        Y1 = torch.empty((1, len(t), len(case.observables)), dtype=torch.float64, device=dev)
        pb.simulate(mean, Y1)
        torch.cuda.synchronize()
        truth = Y1[0].cpu().numpy()
        pop = case.population_ln()
        mu_prior, omega_prior = [(m, 0.5) for m, _ in pop], [0.5] * P
        mu = torch.tensor([[m for m, _ in pop]] * C, dtype=torch.float64, device=dev)
        omega = torch.tensor([[s for _, s in pop]] * C, dtype=torch.float64, device=dev)
        rec = {"chains": C, "timepoints": len(t), "integrator": case.method + " + forward sensitivities",
               "rtol": case.rtol, "atol": case.atol,
               "collective": "1 ncclAllReduce of [8, 5] fp64 per evaluation, on the kernels' stream" if world > 1 else "none (1 rank)"}
        sizes = [("weak", bc.N_INDIVIDUALS_PER_GPU * world, 50)]
        if world > 1:
            sizes.append(("strong_8x1000", 1000, 50))
        sizes.append(("strong_8x16000", 16000, 10))
        for label, n_ind, reps in sizes:
            first, stop = dd.shard_range(n_ind, rank, world)
            L = stop - first
            y, theta = case.inputs(truth, n_ind)
            pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(first, stop)], axis=-1), case.sigma)
            th = torch.from_numpy(np.ascontiguousarray(theta[:, first:stop].transpose(2, 0, 1))).to(dev)     # [P, C, L]
            H = po.HierarchicalObjectiveDevice(pb, C, n_ind, mu_prior, omega_prior, first_individual=first, n_local=L, world=world)
            for _ in range(3):
                H(th, mu, omega)
            ms = _timed(torch, dist, world, lambda: H(th, mu, omega), reps)
            red = H.red.cpu().numpy()
            if not (np.all(np.isfinite(red)) and bool(torch.isfinite(H.dtheta).all()) and int(H.status.abs().sum()) == 0):
                raise SystemExit(f"{case.name} {label}: non-finite logp / gradient or failed trajectories")
            r = {"individuals_total": n_ind, "individuals_this_rank": L, "trajectories_per_eval": C * n_ind,
                 "ms_per_batched_eval": ms, "logp_grad_evals_per_s": C / (ms * 1e-3),
                 "trajectories_per_s": C * n_ind / (ms * 1e-3), "logp_chain0": float(red[0, 0])}
            if label == "weak":
                # the same evaluation through host buffers (what a host-side sampler would pay per step)
                th_pin = torch.empty(th.shape, dtype=torch.float64).pin_memory()
                th_pin.copy_(th)
                out_red = torch.empty(H.red.shape, dtype=torch.float64).pin_memory()
                out_g = torch.empty(H.dtheta.shape, dtype=torch.float64).pin_memory()

                def host_eval():
                    th.copy_(th_pin, non_blocking=True)
                    H(th, mu, omega)
                    out_red.copy_(H.red, non_blocking=True)
                    out_g.copy_(H.dtheta, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                ms_h = _timed(torch, dist, world, host_eval, reps)
                r["e2e"] = {"ms_per_batched_eval": ms_h, "logp_grad_evals_per_s": C / (ms_h * 1e-3),
                            "h2d_bytes_per_eval": int(th.numel() * 8), "d2h_bytes_per_eval": int((H.red.numel() + H.dtheta.numel()) * 8)}
            if label == "weak":
                # the consumers the north star names: gradient-based samplers calling this evaluation at every step, with their own
                # state on the GPU too (a failure here is reported in the record, it does not void the headline line)
                x0 = np.concatenate([np.log(theta).reshape(C, -1), np.array([[m for m, _ in pop]] * C),
                                     np.log(np.array([[s_ for _, s_ in pop]] * C))], axis=1)
                r["hmc_device"] = _sampler_leg(lambda: _hmc_leg(H, x0, world, dist))
                if world == 1:
                    r["nuts_device"] = _sampler_leg(lambda: _nuts_leg(H, x0))
            if label == "strong_8x1000" and world > 1:
                # every rank also evaluates ALL individuals alone: the sharded sums must agree with the unsharded ones
                pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(n_ind)], axis=-1), case.sigma)
                th_all = torch.from_numpy(np.ascontiguousarray(theta.transpose(2, 0, 1))).to(dev)
                H1 = po.HierarchicalObjectiveDevice(pb, C, n_ind, mu_prior, omega_prior)
                ref = H1(th_all, mu, omega)[0].cpu().numpy()
                r["sharded_vs_unsharded_max_rel"] = float(np.max(np.abs(red - ref) / np.abs(ref)))
                if r["sharded_vs_unsharded_max_rel"] > 1e-11:
                    raise SystemExit(f"{case.name}: sharded evaluation differs from the unsharded one: {r}")
            rec[label] = r
        rec["logp_grad_evals_per_s"] = rec["weak"]["logp_grad_evals_per_s"]
        rec["ms_per_batched_eval"] = rec["weak"]["ms_per_batched_eval"]
        out[case.name] = rec
        pb.close()
    out["note"] = ("one batched evaluation = logp + gradient of all 8 chains (each over every individual); logp_grad_evals_per_s "
                   "counts chain evaluations; device resident, CUDA events, max over ranks; settings = bench_cases.py, verified "
                   "against the oracle by tests/test_gpu_parity.py::test_bench_logp_grad_settings_meet_the_bar")
    return out


def extra_measurements(rt, torch, pb, P, logp, status, steps, stream, B, T, th_host) -> dict:
    """SURVEY.md section 8d rows besides the headline: C2 "fused logp, no trajectory store" (device and through host
    buffers) and the C1 single-parameter-set latency (forward simulation and logp through the host API)."""
    out = {}
    for _ in range(3):
        pb.logp(P, logp, status=status, steps=steps, stream=stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 10
    for _ in range(reps):
        pb.logp(P, logp, status=status, steps=steps, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    P_pin = rt.pinned_empty((2, B))
    P_pin[:] = th_host
    lp_pin = rt.pinned_empty((B,))
    for _ in range(2):
        pb.logp_host(P_pin, out=lp_pin)
    t0 = time.perf_counter()
    for _ in range(reps):
        lp_h, _, _, nf = pb.logp_host(P_pin, out=lp_pin)
    dt = (time.perf_counter() - t0) / reps
    out["c2_fused_logp"] = {"what": "C2 with the likelihood folded into the integrator and NO trajectory store (pv_logp)",
                            "trajectories_per_s": B / (ms * 1e-3), "ms_per_step": ms,
                            "e2e_trajectories_per_s": B / dt, "e2e_ms_per_step": dt * 1e3, "e2e_api": "pv_logp_host",
                            "h2d_bytes_per_step": int(16 * B), "d2h_bytes_per_step": int(8 * B),
                            "e2e_matches_device": bool(np.array_equal(lp_h[:1000], logp[:1000].cpu().numpy()))}      # same batch, same warps: bit-identical
    # C1: one parameter set (defaults), the reference's own CPU-runnable case: latency of one call through the host API
    c1 = {}
    for Tn in (21, 50):
        p1 = rt.Problem("simple_pk", device=pb.device)
        p1.set_solver(rtol=RTOL, atol=ATOL)
        p1.set_columns([])
        p1.set_timepoints(np.linspace(0.0, 100.0, Tn))
        p1.set_observables(OBS)
        Y1, _ = p1.simulate_host(None, 1)
        d = c1_data(Y1[0])
        p1.set_data(np.stack([np.ones_like(d), d, d * d])[..., None], 1.0, rt.NOISE_NORMAL)
        for _ in range(20):
            p1.simulate_host(None, 1)
        n = 300
        t0 = time.perf_counter()
        for _ in range(n):
            p1.simulate_host(None, 1)
        us_sim = (time.perf_counter() - t0) / n * 1e6
        for _ in range(20):
            p1.logp_host(np.empty((0, 1)))
        t0 = time.perf_counter()
        for _ in range(n):
            p1.logp_host(np.empty((0, 1)))
        us_lp = (time.perf_counter() - t0) / n * 1e6
        c1[f"T{Tn}"] = {"simulate_host_us": us_sim, "logp_host_us": us_lp}
        p1.close()
    c1["what"] = "C1: simple_pk, one parameter set, host call -> result on the host (launch + integration + copies), microseconds"
    out["c1_latency"] = c1
    return out


# ----------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=400000)
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU (default = the named config)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    rt = importlib.import_module(PKG + ".runtime")
    lib = rt.load_library()
    if lib.pv_device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: parvar_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    # one process per GPU: run next to it (pinned pages land on the NUMA node of the thread that touches them first)
    importlib.import_module(PKG + ".numa").bind_to_gpu_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch
    T = T_POINTS
    tgrid = np.linspace(0.0, 100.0, T)

    pb = rt.Problem("simple_pk", device=local)
    pb.set_solver(method=rt.METHOD_AUTO, rtol=RTOL, atol=ATOL)   # the default path: DOPRI5 + RODAS4 re-run of stiff rows
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(tgrid)
    pb.set_observables(OBS)
    Y1 = torch.empty((1, T, 3), dtype=torch.float64, device=dev)
    pb.simulate(torch.ones((2, 1), dtype=torch.float64, device=dev), Y1)       # defaults: k_abs = CL = 1
    torch.cuda.synchronize()
    data = c1_data(Y1[0].cpu().numpy())
    stats = np.stack([np.ones_like(data), data, data * data])[..., None]
    pb.set_data(stats, 1.0, rt.NOISE_NORMAL)

    th_host = draws(rank, B)
    P = torch.from_numpy(th_host).to(dev)
    Y = torch.empty((B, T, 3), dtype=torch.float64, device=dev)
    logp = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    steps = torch.empty((2, B), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also counts steps for the algorithmic-flop figure; deterministic across launches) -------------
    for _ in range(max(args.warmup, 1)):
        pb.simulate_logp(P, Y, logp, status=status, steps=steps, stream=stream)
    torch.cuda.synchronize()
    n_failed = int((status != 0).sum())
    if n_failed:
        raise SystemExit(f"{n_failed} trajectories failed to integrate")
    executed = int(steps.sum())               # steps the launch executes (lockstep warps: every lane takes its warp's steps)
    # ALGORITHMIC steps = what each trajectory needs on its own (per-lane adaptive stepping): counted once with the
    # lockstep schedule switched off; the roofline's flops use this figure, not the (slightly larger) executed one
    pb.set_lockstep(1)
    pb.simulate_logp(P, Y, logp, status=status, steps=steps, stream=stream)
    torch.cuda.synchronize()
    pb.set_lockstep(0)
    attempted = int(steps.sum())
    accepted = int(steps[0].sum())
    steps_max = int(steps.sum(dim=0).max())
    pb.simulate_logp(P, Y, logp, status=status, steps=steps, stream=stream)     # back to the default schedule
    torch.cuda.synchronize()
    fl_step = pb.flops("step_dopri5")
    fl_dense = pb.flops("dense_dopri5")
    fl_hoist = pb.flops("hoist")
    fl_logp = 8 * T * 3
    flops_per_launch = attempted * fl_step + B * (T * fl_dense + fl_hoist + fl_logp + 6 * pb.flops("rhs"))
    bytes_per_launch = B * (8 * 2 + 8 * T * 3 + 8 + 4 + 8)

    # ---- timed region: value ------------------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        # nvidia-smi needs ~0.3 s to start reporting: keep the GPU under the same load until it does
        t_spin = time.perf_counter()
        while time.perf_counter() - t_spin < 0.5:
            pb.simulate_logp(P, Y, logp, status=status, steps=steps, stream=stream)
            torch.cuda.synchronize()
    n0 = rt.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        pb.simulate_logp(P, Y, logp, status=status, steps=steps, stream=stream)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = rt.launch_count() - n0
    clk = clocks.stop() if rank == 0 else {}
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * args.steps / (ms_max * 1e-3)
    checksum = float(logp.sum().item())

    # ---- e2e: host buffers through the C ABI --------------------------------------------------------------
    P_pin = rt.pinned_empty((2, B))
    P_pin[:] = th_host
    Y_pin = rt.pinned_empty((B, T, 3))
    lp_pin = rt.pinned_empty((B,))
    for _ in range(2):
        pb.simulate_host(P_pin, B, Y=Y_pin, logp=lp_pin)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, nf = pb.simulate_host(P_pin, B, Y=Y_pin, logp=lp_pin)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t_e.item())
    # the host pipeline integrates 32 768-row chunks, the device call the whole batch: different warps under the lockstep
    # schedule, so the two agree within the integration tolerance, not bit for bit
    e2e_ok = bool(np.allclose(lp_pin[:1000], logp[:1000].cpu().numpy(), rtol=1e-9, atol=0))

    lg = logp_grad_leg(rt, torch, dist, rank, world, local)
    extra = extra_measurements(rt, torch, pb, P, logp, status, steps, stream, B, T, th_host) if rank == 0 else {}
    if world > 1:
        dist.barrier()

    if rank == 0:
        peak64 = rt.fma_peak_tflops(local, fp64=True)
        peak32 = rt.fma_peak_tflops(local, fp64=False)
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if peaks_file.exists():
            hbm_peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
            hbm_src = "measured (MEASURED_PEAKS.json)"
        traffic = None
        tf = ROOT / "profiles" / "r02_dominant_kernel_traffic.json"
        if tf.exists() and B == B_PER_GPU:
            traffic = json.loads(tf.read_text())["traffic_bytes_per_launch"]     # from the committed ncu --set full capture
        kernel_ms = ms / args.steps
        ach_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12
        ach_gbs = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        cpu = cpu_baseline(args.cpu_sample)
        line = {
            "metric": "ODE trajectories/s (simple_pk forward simulation + logp)",
            "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int(2 * B * 8),
                    "d2h_bytes_per_step": int(T * 3 * B * 8 + B * 8), "steps": e2e_steps,
                    "api": "pv_simulate_host (C ABI, pinned host buffers, chunked H2D/kernel/D2H overlap)",
                    "matches_device_path": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "fp64_fma", "achieved": ach_tf, "peak": peak64, "unit": "TFLOP/s",
                         "frac": ach_tf / peak64 if peak64 > 0 else None, "traffic": traffic,
                         "traffic_note": "dram read+write bytes per launch of the dominant kernel (ncu --set full, profiles/); algorithmic "
                                         f"bytes per launch = {bytes_per_launch}",
                         "peak_source": "measured in this run: pv_fma_peak_tflops (register-resident DFMA chains)",
                         "fp32_fma_peak_tflops": peak32,
                         "kernel": "pv_batch_lockstep_kernel<pv_model_simple_pk, TRAJ|LOGP>",
                         "kernel_ms": kernel_ms,
                         "kernel_ms_note": "timed region / steps, i.e. the whole step: the lockstep kernel is 96.9 % of it in the ncu "
                                           "launch list (profiles/r02_bench_launch_list_staged.csv: 2.47 of 2.55 ms; the rest is the "
                                           "locality sort's seven kernels and METHOD_AUTO's compaction + empty RODAS4 pass)",
                         "flops_per_launch": flops_per_launch,
                         "flops_per_attempted_step": fl_step,
                         "attempted_steps_per_trajectory": attempted / B, "accepted_steps_per_trajectory": accepted / B,
                         "executed_steps_per_trajectory": executed / B,
                         "flops_note": "algorithmic flops = steps of per-trajectory adaptive stepping x generator flop counts; the default "
                                       "schedule (locality sort + lockstep warps) executes executed_steps_per_trajectory",
                         "max_steps_in_batch": steps_max,
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "peak_source": hbm_src, "bytes_per_launch": bytes_per_launch}},
            "cpu_baseline": cpu,
            "logp_grad": lg,
            **extra,
            "logp_checksum": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
