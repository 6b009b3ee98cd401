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
            "integrator": "METHOD_AUTO: DOPRI5 adaptive with dense output; rows that exhaust 4000 explicit steps are re-run with RODAS4 (none in this cloud)", "rtol": RTOL, "atol": ATOL,
            "l2": "outputs (1.2 GB per step) exceed the 126 MB L2, so every step streams to HBM",
            "sharding": f"independent draws, {n_gpus} rank(s), no data-path collective"}


# ----------------------------------------------------------------------------------------------------------
def logp_grad_leg(rt, torch, dist, rank: int, world: int, local: int, evals: int = 20) -> dict:
    """Second half of BASELINE.json's metric: logp + gradient evaluations/s (configs[2] and [3]).

    Hierarchical objective over 8 chains x (1000 individuals per GPU), individuals sharded over the ranks (weak
    scaling), ONE all-reduce of [chains, 1 + 2P] per evaluation; per-individual gradients stay on their rank.
    Each evaluation goes through the public host API (`HierarchicalObjective.local_terms` -> `pv_logp_grad_host`):
    host theta -> H2D -> sensitivity kernel -> D2H of logp / gradient -> population terms -> all-reduce.  Data are
    synthetic: the model's default trajectory (through the CUDA path) x (1 + 0.05 eps) per individual.
    """
    po = importlib.import_module(PKG + ".optimization.petab_optimization")
    dd = importlib.import_module(PKG + ".distributed")
    pf = importlib.import_module(PKG + ".experiments.petab_factory")
    C, n_loc, T = 8, 1000, 20
    n_ind = n_loc * world
    first, stop = dd.shard_range(n_ind, rank, world)
    t = np.linspace(0.0, 100.0, T)
    out = {}
    cases = [("C3 simple_pk", "simple_pk", ["k_abs", "CL"], OBS, {}, (1.0, 1.0), (1.0, 1.0), 0.05, (1e-8, 1e-11), "DOPRI5"),
             ("C4 icg_body_flat", "icg_body_flat", ["BW", "LI__ICGIM_Vmax"], ["Cre_plasma_icg", "Cgi_plasma_icg"],
              {"IVDOSE_icg": 10.0}, (75.0, 10.0), (0.0369598840327503, 0.01), 1e-4, (1e-7, 1e-11), "RODAS4")]
    for name, model, cols, obs, fixed, ms1, ms2, sigma, (rtol, atol), method in cases:
        pb = rt.Problem(model, device=local)
        pb.set_solver(rtol=rtol, atol=atol)
        for k, v in fixed.items():
            pb.set_value(k, v)
        pb.set_columns(cols)
        pb.set_timepoints(t)
        pb.set_observables(obs)
        mean = torch.tensor([[ms1[0]], [ms2[0]]], dtype=torch.float64, device="cuda")
        Y1 = torch.empty((1, T, len(obs)), dtype=torch.float64, device="cuda")
        pb.simulate(mean, Y1)
        torch.cuda.synchronize()
        truth = Y1[0].cpu().numpy()
        rs = np.random.RandomState(1234)                 # every rank draws the full set and keeps its slice
        y = (truth[None] * (1 + 0.05 * rs.normal(size=(n_ind,) + truth.shape)))[first:stop]
        (m1, s1), (m2, s2) = pf.lognorm_par_transform(*ms1), pf.lognorm_par_transform(*ms2)
        theta = np.stack([rs.lognormal(m1, s1, size=(C, n_ind)), rs.lognormal(m2, s2, size=(C, n_ind))], axis=2)
        theta = np.ascontiguousarray(theta[:, first:stop])
        pb.set_data(np.stack([po.sufficient_statistics(y[i:i + 1], 0) for i in range(len(y))], axis=-1), sigma)
        H = po.HierarchicalObjective(pb, n_ind, mu_prior=[(m1, 0.5), (m2, 0.5)], omega_prior=[0.5, 0.5],
                                     first_individual=first, n_local=stop - first)
        mu = np.tile(np.array([m1, m2]), (C, 1))
        omega = np.tile(np.array([s1, s2]), (C, 1))

        def evaluate():
            lp, dth, dmu, dom = H.local_terms(theta, mu, omega)
            buf = np.ascontiguousarray(np.concatenate([lp[:, None], dmu, dom], axis=1))          # [C, 1 + 2P]
            if world > 1:
                tb = torch.from_numpy(buf).cuda()
                dist.all_reduce(tb, op=dist.ReduceOp.SUM)
                buf = tb.cpu().numpy()
            return H.finish(buf[:, 0], buf[:, 1:3], buf[:, 3:5], mu, omega) + (dth,)

        for _ in range(3):
            res = evaluate()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(evals):
            res = evaluate()
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        if not (np.all(np.isfinite(res[0])) and np.all(np.isfinite(res[3]))):
            raise SystemExit(f"{name}: non-finite logp / gradient")
        out[name] = {"logp_grad_evals_per_s": C * evals / dt, "ms_per_batched_eval": 1e3 * dt / evals,
                     "kernel_ms": pb.last_kernel_ms, "trajectories_per_eval": C * n_ind,
                     "trajectories_per_s": C * n_ind * evals / dt, "chains": C,
                     "individuals_per_gpu": n_loc, "timepoints": T, "integrator": method + " + forward sensitivities",
                     "rtol": rtol, "atol": atol, "collective": "1 all-reduce of [8, 5] fp64 per evaluation" if world > 1 else "none (1 rank)",
                     "logp_chain0": float(res[0][0])}
        pb.close()
    out["note"] = ("one batched evaluation = logp + gradient of all 8 chains (each over every individual); "
                   "logp_grad_evals_per_s counts chain evaluations; timed through the host API, max over ranks")
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
    attempted = int(steps.sum())
    accepted = int(steps[0].sum())
    steps_max = int(steps.sum(dim=0).max())
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
    e2e_ok = bool(np.allclose(lp_pin[:1000], logp[:1000].cpu().numpy(), rtol=0, atol=0))

    lg = logp_grad_leg(rt, torch, dist, rank, world, local)

    if rank == 0:
        peak64 = rt.fma_peak_tflops(local, fp64=True)
        peak32 = rt.fma_peak_tflops(local, fp64=False)
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if peaks_file.exists():
            hbm_peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
            hbm_src = "measured (MEASURED_PEAKS.json)"
        traffic = None
        tf = ROOT / "profiles" / "r01_dominant_kernel_traffic.json"
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
                         "kernel": "pv_batch_kernel<pv_model_simple_pk, DOPRI5, TRAJ|LOGP>",
                         "kernel_ms": kernel_ms, "flops_per_launch": flops_per_launch,
                         "flops_per_attempted_step": fl_step,
                         "attempted_steps_per_trajectory": attempted / B, "accepted_steps_per_trajectory": accepted / B,
                         "max_steps_in_batch": steps_max,
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "peak_source": hbm_src, "bytes_per_launch": bytes_per_launch}},
            "cpu_baseline": cpu,
            "logp_grad": lg,
            "logp_checksum": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
