"""GPU mirror of the reference's estimation driver.

Reference: ``PyPestoSampler`` (load_problem / run_optimization / bayesian_sampler / results_dict / results_df,
src/parvar/optimization/petab_optimization.py:29-269) and ``optimize_experiment`` / ``optimize_experiments``
(:272-368).  Same stages, same settings dictionary, same ``<uid>_results.tsv`` columns, so the reference's analysis
(``src/parvar/analysis/utils.py:28-98``) reads the output unchanged - but all experiments that share a model and a
measurement grid are advanced TOGETHER (one launch per optimiser iteration / sampler iteration for all of them),
instead of one process per experiment (``optimize_experiments_multicore``, :370-391).
"""
from __future__ import annotations

import json
import time
import traceback
from pathlib import Path
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from . import optimizer as opt_mod
from . import sampler as smp_mod
from .petab_io import load_petab_problem, parse_petab_problem, read_petab_tables
from .petab_optimization import GroupData, PetabObjectiveBatch

DEFAULT_SETTINGS = dict(          # petab_optimization.py:294-306
    maxiter=1e4, fatol=1e-12, frtol=1e-12, optim="fides", startpoint_method="uniform", n_starts=100, seed=1234,
    create_optimization_plots=False, engine="gpu-batched", n_samples=5000, n_chains=4,
)


def get_group_from_pid(pid: str) -> Optional[str]:
    """``k_abs_MALE`` -> ``MALE`` (reference src/parvar/experiments/utils.py:48-62)."""
    last = pid.rsplit("_", 1)[-1]
    return last if last and last.isascii() else None


def get_parameter_from_pid(pid: str) -> str:
    """``k_abs_MALE`` -> ``k_abs`` (reference src/parvar/experiments/utils.py:65-87)."""
    idx = pid.rfind("_")
    if idx == -1:
        return pid
    last = pid[idx + 1:]
    return pid[:idx] if last and last.isascii() else pid


class GpuPetabSampler:
    """Optimise + sample Q PEtab problems of one model together.  Mirrors ``PyPestoSampler`` (:29-269)."""

    def __init__(self, objective: PetabObjectiveBatch, uids: Optional[Sequence[str]] = None,
                 start_box: Optional[tuple[np.ndarray, np.ndarray]] = None):
        self.objective = objective
        self.uids = list(uids) if uids is not None else [f"problem{q}" for q in range(objective.Q)]
        self.start_box = start_box
        self.optimize_result: Optional[opt_mod.OptimizeResult] = None
        self.sample_result: Optional[smp_mod.SampleResult] = None
        self.optim_duration: Optional[float] = None
        self.sampler_duration: Optional[float] = None

    # ---- objective adapters ------------------------------------------------------------------------------------------
    def _eval_opt(self, X):
        o = self.objective.evaluate(X, order=2)
        fval = -(o["llh"] + o["lprior"])
        grad = -(o["dllh"] + o["dlprior"])
        dim = X.shape[-1]
        hess = o["fim"] + o["prior_precision"][..., None] * np.eye(dim)
        bad = ~np.isfinite(fval)
        fval = np.where(bad, np.inf, fval)
        return {"fval": fval, "grad": grad, "hess": hess}

    def _eval_sample(self, X):
        o = self.objective.evaluate(X, order=0)
        return o["llh"], o["lprior"]

    # ---- stages ------------------------------------------------------------------------------------------------------------
    def run_optimization(self, maxiter: float = 1e4, fatol: float = 1e-12, frtol: float = 1e-12, n_starts: int = 100,
                         seed: int = 1234, device_loop: bool = True, **_):
        """Multistart optimisation (:99-156).  Start points are uniform in the bounds like the reference
        (``pypesto.startpoint.uniform`` on [0, 1e8], :120-121) unless a ``start_box`` was given."""
        lb, ub = self.objective.lb, self.objective.ub
        slb, sub = self.start_box if self.start_box is not None else (lb, ub)
        x0 = opt_mod.uniform_startpoints(n_starts, np.asarray(slb, float), np.asarray(sub, float), Q=self.objective.Q, seed=seed)
        t0 = time.perf_counter()
        if device_loop:       # the whole loop on the GPU; False = numpy loop, one objective launch per iteration
            self.optimize_result = opt_mod.minimize_device(self.objective, x0, maxiter=int(min(maxiter, 1e4)), fatol=fatol, frtol=frtol)
        else:
            self.optimize_result = opt_mod.minimize(self._eval_opt, x0, lb, ub, maxiter=int(min(maxiter, 1e4)), fatol=fatol, frtol=frtol)
        self.optim_duration = time.perf_counter() - t0
        return self.optimize_result

    def bayesian_sampler(self, n_samples: int = 5000, n_chains: int = 4, seed: int = 1234, device_loop: bool = True):
        """Adaptive Metropolis inside adaptive parallel tempering, started at the best optimisation result (:158-194).
        ``device_loop``: the whole loop on the GPU (``sampler.sample_device``); False = host loop, one objective launch per
        iteration (``sampler.sample``) - same random stream, same chains up to round-off."""
        if self.optimize_result is None:
            raise RuntimeError("run_optimization first (pyPESTO starts the chains at the best optimisation result)")
        x0 = self.optimize_result.x[:, 0, :]
        t0 = time.perf_counter()
        if device_loop:
            self.sample_result = smp_mod.sample_device(self.objective, x0, n_samples=n_samples, n_chains=n_chains, seed=seed)
        else:
            self.sample_result = smp_mod.sample(self._eval_sample, x0, self.objective.lb, self.objective.ub,
                                                n_samples=n_samples, n_chains=n_chains, seed=seed)
        t1 = time.perf_counter()
        smp_mod.effective_sample_size(self.sample_result)
        self.sampler_duration = time.perf_counter() - t0
        self.ess_duration = time.perf_counter() - t1
        return self.sample_result

    # ---- results (columns of results_df, :254-269) ---------------------------------------------------------------------------
    def results_dict(self, q: int, settings: dict, pool_tempered_chains: bool = True) -> dict:
        """Posterior statistics of problem q.  The reference feeds ``trace_x`` of ALL tempered chains to arviz
        (``az.convert_to_inference_data(trace_)``, :204-211), so its mean / std / median / HDI pool the hot chains as
        well; ``pool_tempered_chains=True`` reproduces that, ``False`` uses the beta = 1 chain only."""
        tr = self.sample_result.trace_x[q]                    # [C, n+1, dim]
        vals = tr if pool_tempered_chains else tr[:1]
        res = {}
        for i, pid in enumerate(self.objective.x_names):
            v = vals[:, :, i]
            lo, hi = smp_mod.hdi(v)
            res[pid] = {"mean": float(np.mean(v)), "std": float(np.std(v)), "median": float(np.median(v)),
                        "ess": float(self.sample_result.effective_sample_size[q]), "n_samples": settings["n_samples"],
                        "optim_duration": self.optim_duration, "sampler_duration": self.sampler_duration,
                        "values": v, "hdi_low": lo, "hdi_high": hi}
        return res

    def results_table(self, settings: dict, pool_tempered_chains: bool = True) -> pd.DataFrame:
        """``results_df`` of ALL problems of the batch in one frame, without the ``values`` column: the same statistics
        (mean / std / median / 94 % HDI over the pooled chains, ESS) computed for every (problem, parameter) at once - one
        sort per batch instead of two per parameter and one small DataFrame per problem (at 2160 problems the per-problem
        frames were half of the study's estimation time, and they hold the interpreter lock).  Rows in ``results_df`` order."""
        tr = self.sample_result.trace_x                       # [Q, C, n+1, dim]
        vals = tr if pool_tempered_chains else tr[:, :1]
        Q, _, _, dim = vals.shape
        v = np.ascontiguousarray(np.moveaxis(vals, 3, 1)).reshape(Q, dim, -1)      # [Q, dim, C (n+1)]
        n = v.shape[2]
        mean, std = v.mean(axis=2), v.std(axis=2)
        v.sort(axis=2)                                        # in place: v is this function's copy
        median = 0.5 * (v[:, :, (n - 1) // 2] + v[:, :, n // 2])
        k = int(np.floor(0.94 * n))                           # smp_mod.hdi's default probability
        if k < 1 or k >= n:
            lo, hi = v[:, :, 0], v[:, :, -1]
        else:
            i = np.argmin(v[:, :, k:] - v[:, :, : n - k], axis=2)[..., None]
            lo = np.take_along_axis(v, i, axis=2)[..., 0]
            hi = np.take_along_axis(v, i + k, axis=2)[..., 0]
        names = list(self.objective.x_names)
        ess = np.asarray(self.sample_result.effective_sample_size, dtype=np.float64)
        return pd.DataFrame({
            "id": np.repeat(np.asarray(self.uids, dtype=object), dim),
            "group": [get_group_from_pid(pid) for pid in names] * Q,
            "parameter": [get_parameter_from_pid(pid) for pid in names] * Q,
            "pid": names * Q,
            "mean": mean.ravel(), "std": std.ravel(), "median": median.ravel(), "ess": np.repeat(ess, dim),
            "n_samples": settings["n_samples"], "optim_duration": self.optim_duration, "sampler_duration": self.sampler_duration,
            "hdi_low": lo.ravel(), "hdi_high": hi.ravel()})

    def results_df(self, q: int, settings: dict, pool_tempered_chains: bool = True) -> pd.DataFrame:
        rows = []
        for pid, stats in self.results_dict(q, settings, pool_tempered_chains).items():
            rows.append({"id": self.uids[q], "group": get_group_from_pid(pid), "parameter": get_parameter_from_pid(pid),
                         "pid": pid, **stats})
        return pd.DataFrame(rows)


def _batch_from_yaml(yaml_paths: Sequence[Path], mode: str, dosage, device: int) -> PetabObjectiveBatch:
    """One batched objective for PEtab directories that share model, parameters, observables, time grid, noise formula,
    bounds and (literal mode) condition-table initial values; measurements and priors are per problem.  The tables are
    parsed without creating GPU objects; ONE ``Problem`` is built for the whole batch."""
    parsed = [parse_petab_problem(p) for p in yaml_paths]
    first = parsed[0]
    for p, o in zip(yaml_paths, parsed):
        same = (o.model, o.parameters, o.species) == (first.model, first.parameters, first.species) and \
            np.array_equal(o.timepoints, first.timepoints) and np.array_equal(o.sigma, first.sigma) and \
            np.array_equal(o.lb, first.lb) and np.array_equal(o.ub, first.ub) and o.literal_values == first.literal_values
        if not same:
            raise ValueError(f"{p}: experiments batched together must share model, parameters, observables, time grid, noise "
                             "formula, bounds and condition-table initial values")
    batch = PetabObjectiveBatch(first.model, first.parameters, [o.groups for o in parsed], first.timepoints, first.species,
                                sigma=first.sigma, priors=[o.priors for o in parsed], mode=mode, dosage=dosage, device=device,
                                initial_values=first.literal_values if mode == "literal" else None)
    batch.lb, batch.ub = first.lb, first.ub
    return batch


def optimize_experiments(results_dir: Path, yaml_paths: Sequence[Path], caching: bool = True, settings: Optional[dict] = None,
                         mode: str = "literal", dosage: Optional[dict] = None, device: int = 0,
                         start_box: Optional[tuple[np.ndarray, np.ndarray]] = None) -> list[bool]:
    """Reference ``optimize_experiments`` (:351-368) for experiments that share model and time grid: optimise and
    sample them together, write one ``<uid>_results.tsv`` per experiment (errors -> ``<uid>_errors.tsv``, :340-348)."""
    results_dir = Path(results_dir)
    results_dir.mkdir(parents=True, exist_ok=True)
    settings = {**DEFAULT_SETTINGS, **(settings or {})}
    todo = [Path(p) for p in yaml_paths if not (caching and (results_dir / f"{Path(p).parent.name}_results.tsv").exists())]
    ok = {Path(p): True for p in yaml_paths}
    if not todo:
        return [True] * len(yaml_paths)
    uids = [p.parent.name for p in todo]
    try:
        batch = _batch_from_yaml(todo, mode, dosage, device)
        drv = GpuPetabSampler(batch, uids=uids, start_box=start_box)
        drv.run_optimization(maxiter=settings["maxiter"], fatol=settings["fatol"], frtol=settings["frtol"],
                             n_starts=settings["n_starts"], seed=settings["seed"])
        drv.bayesian_sampler(n_samples=settings["n_samples"], n_chains=settings["n_chains"], seed=settings["seed"])
        for q, uid in enumerate(uids):
            df = drv.results_df(q, settings)
            df["values"] = df["values"].apply(lambda x: json.dumps(np.asarray(x).tolist()))
            df.to_csv(results_dir / f"{uid}_results.tsv", sep="\t", index=False)
    except Exception:
        err = traceback.format_exc()
        for p, uid in zip(todo, uids):
            (results_dir / f"{uid}_errors.tsv").write_text(err)
            ok[p] = False
    return [ok[Path(p)] for p in yaml_paths]


def optimize_experiment(yaml_path: Path, results_dir: Path, caching: bool = True, **kwargs) -> bool:
    """Reference ``optimize_experiment`` (:272-348) for a single PEtab problem."""
    return optimize_experiments(results_dir, [Path(yaml_path)], caching=caching, **kwargs)[0]


def definitions_dataframe(experiments: Sequence[dict]) -> pd.DataFrame:
    """``definitions.tsv`` in the reference's schema: ``PETabExperimentList.to_dataframe`` (src/parvar/experiments/
    experiment.py:170-204) with the timepoints column as ``create_petabs`` stores it (steps + 1, petab_factory.py:541-543).

    ``experiments``: dicts with id, model, prior_type, samples, timepoints, noise_cv and
    ``sampling = {group: {parameter: (loc, scale)}}``."""
    rows = []
    for xp in experiments:
        for group, pars in xp["sampling"].items():
            for par, (loc, scale) in pars.items():
                rows.append({"id": xp["id"], "model": xp["model"], "prior_type": xp["prior_type"], "n_groups": len(xp["sampling"]),
                             "group": group, "samples": xp["samples"], "timepoints": xp["timepoints"], "noise_cv": xp["noise_cv"],
                             "parameter": par, "dsn_type": "DistributionType.LOGNORMAL",
                             "dsn_par": str({"loc": float(loc), "scale": float(scale)})})     # pydantic stores floats
    return pd.DataFrame(rows)


def write_reference_layout(results_path: Path, xp_type: str, experiments: Sequence[dict], result_frames: dict[str, pd.DataFrame]) -> Path:
    """Write what the reference's analysis reads (``join_optimization_results``, src/parvar/analysis/utils.py:28-98):
    ``<results_path>/xps/<xp_type>/definitions.tsv`` and ``.../optimization_results/<uid>_results.tsv`` (one per
    experiment, ``values`` as a JSON list, petab_optimization.py:334-336)."""
    base = Path(results_path) / "xps" / xp_type
    (base / "optimization_results").mkdir(parents=True, exist_ok=True)
    definitions_dataframe(experiments).to_csv(base / "definitions.tsv", sep="\t", index=False)
    for uid, df in result_frames.items():
        df = df.copy()
        df["values"] = df["values"].apply(lambda x: x if isinstance(x, str) else json.dumps(np.asarray(x).tolist()))
        df.to_csv(base / "optimization_results" / f"{uid}_results.tsv", sep="\t", index=False)
    return base
