"""Host-side mirror of the hot-path part of the reference's ``petab_factory`` module.

Same names, argument meaning and error behaviour as reference
``src/parvar/experiments/petab_factory.py`` for

  * ``lognorm_par_transform``            (:94-98)
  * ``create_samples_parameters``        (:101-121)   numpy legacy global RNG, incl. the loc/scale swap
  * ``SimulationSettings``               (:80-91)
  * ``ODESampleSimulator``               (:180-272)   -> ONE batched CUDA launch instead of a Python
                                                        loop of roadrunner calls

so that ``create_petab_for_experiment`` (:548-621) works unchanged when it imports these names
from here.  The PEtab writer, plots and the sweep factory are out of scope.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from enum import Enum
from pathlib import Path
from typing import Any, Hashable, Optional

import numpy as np
import pandas as pd

from .. import runtime
from .dataset import SimDataset
from .experiment import DistributionType, Noise, Observable

try:  # optional: return a real xr.Dataset when xarray is installed (it is not in this image)
    import xarray as _xr
except Exception:  # pragma: no cover
    _xr = None


class Category(str, Enum):
    """Categories (reference :37-43)."""
    MALE = "male"
    FEMALE = "female"
    OLD = "old"
    YOUNG = "young"


class PKPDParameters(str, Enum):
    """Sampled parameters of the three models (reference :54-68)."""
    k1 = "k1"
    CL = "CL"
    k_abs = "k_abs"
    LI__ICGIM_Vmax = "LI__ICGIM_Vmax"
    BW = "BW"


@dataclass
class LognormParameters:
    """Reference :71-77."""
    n: int        # number of samples
    sigma: float  # standard deviation
    mu: float     # mean


@dataclass
class SimulationSettings:
    """Reference :80-91."""
    start: float
    end: float
    steps: int
    noise: Any
    observables: Optional[list]
    dosage: Optional[dict[str, float]] = None


def lognorm_par_transform(loc: float, scale: float) -> tuple[float, float]:
    """(mean, sd) -> (mu_ln, sigma_ln) of the log-normal.  Reference :94-98."""
    sigma_ln = np.sqrt(np.log(1 + (scale / loc) ** 2))
    mu_ln = np.log(loc) - sigma_ln ** 2 / 2
    return mu_ln, sigma_ln


def create_samples_parameters(parameters: dict[Hashable, dict[Hashable, LognormParameters]],
                              seed: Optional[int] = 1234,
                              literal: bool = True) -> dict[Hashable, dict[Hashable, np.ndarray]]:
    """Log-normal parameter samples per (group, parameter).  Reference :101-121.

    ``literal=True`` reproduces the reference bit for bit: it seeds numpy's *global* legacy
    generator once and passes ``(sigma, mu)`` to ``lognorm_par_transform(loc, scale)`` - i.e. with
    loc and scale swapped (:114-117).  ``literal=False`` gives the intended
    ``(loc=mu, scale=sigma)`` parameterisation on the same random stream.
    """
    if seed is not None:
        np.random.seed(seed)
    samples: dict[Hashable, dict[Hashable, np.ndarray]] = {}
    for category, pars in parameters.items():
        per_par: dict[Hashable, np.ndarray] = {}
        for par, ln in pars.items():
            if literal:
                mu_ln, sigma_ln = lognorm_par_transform(ln.sigma, ln.mu)
            else:
                mu_ln, sigma_ln = lognorm_par_transform(ln.mu, ln.sigma)
            per_par[par] = np.random.lognormal(mu_ln, sigma_ln, size=ln.n)
        samples[category] = per_par
    return samples


# ----------------------------------------------------------------------------------------
_FINGERPRINTS: Optional[dict[str, str]] = None


def _builtin_fingerprints() -> dict[str, str]:
    global _FINGERPRINTS
    if _FINGERPRINTS is None:
        gen = Path(__file__).resolve().parent.parent / "csrc" / "generated"
        _FINGERPRINTS = {}
        for f in gen.glob("*.json"):
            meta = json.loads(f.read_text())
            _FINGERPRINTS[meta["fingerprint"]] = meta["name"]
    return _FINGERPRINTS


def resolve_model(model_path: Path | str, sens: Optional[list[str]] = None, compile_unknown: bool = True) -> str:
    """Map an SBML file (the reference's annotated file or the math-only copy) to the compiled
    model it describes, by the fingerprint of its ODE system - not by file name.  A file outside the registry - or a
    registry model asked for another sensitivity-parameter set ``sens`` - is generated, compiled (nvcc) and cached as its
    own library on first use, like the reference's roadrunner JIT / AMICI model build (:183-190; run_optimization.py:43-46)."""
    from .. import codegen
    model_path = Path(model_path)
    if not model_path.exists():
        raise FileNotFoundError(model_path)
    sb = codegen.read_sbml(model_path)
    ode = codegen.build_ode_system(sb)
    ode.name = model_path.stem
    for name, builtin_sens in codegen.BUILTIN.items():
        fp = codegen.model_fingerprint(ode, builtin_sens)
        if _builtin_fingerprints().get(fp) == name and (sens is None or set(sens) <= set(builtin_sens)):
            return name
    if not compile_unknown:
        raise runtime.ParvarB200Error(
            f"{model_path}: this ODE system is not compiled into libparvar_b200.so "
            f"(compiled: {sorted(set(_builtin_fingerprints().values()))})")
    name, lib_path = codegen.compile_model(model_path, sens or [])
    name = f"{name}@{lib_path.stem.rsplit('_', 1)[-1]}"      # registry key: model + fingerprint (several builds may coexist)
    if name not in runtime._model_libs:
        runtime.load_model_library(name, lib_path)
    return name


class ODESampleSimulator:
    """Performs simulations with given model and samples - batched on the GPU.

    Reference :180-272.  ``abs_tol`` / ``rel_tol`` are the integrator tolerances as in the
    reference (:183-189), but the defaults are tighter (1e-10 / 1e-7 instead of 1e-6 / 1e-6,
    quirk Q9) so that trajectories agree with a tight-tolerance CPU solve within rtol 1e-6 /
    atol 1e-9; pass the reference's values to reproduce its accuracy.
    """

    def __init__(self, model_path: Path, abs_tol: float = 1e-10, rel_tol: float = 1e-7, device: int = 0,
                 method: int = runtime.METHOD_AUTO, max_steps: int = 20000):
        # any SBML file, as the reference's RoadRunner(str(model_path)) (:185): a registry model by fingerprint, anything
        # else generated + compiled + cached on first use
        self.model_name = resolve_model(model_path)
        self.problem = runtime.Problem(self.model_name, device=device)
        # 20000 = roadrunner's CVODE default `maximum_num_steps`; rows that need more raise like the reference does
        self.problem.set_solver(method=method, rtol=rel_tol, atol=abs_tol, max_steps=max_steps)
        # roadrunner getIds() analogue (:190): concentrations, parameters
        self.ids: list[str] = [f"[{s}]" for s in self.problem.state_ids] + list(self.problem.theta_ids)
        self.last_status: Optional[np.ndarray] = None
        self.last_steps: Optional[np.ndarray] = None

    @staticmethod
    def add_noise(df_sim: pd.DataFrame, coef_variation: float, dsn_type: Any = DistributionType.NORMAL,
                  seed: Optional[int] = None) -> pd.DataFrame:
        """Proportional Gaussian noise, reseeding the global RNG per call.  Reference :192-219."""
        if seed is None:
            seed = np.random.randint(low=0, high=2001)
        np.random.seed(seed)
        df_sim = df_sim.reset_index()
        cols = [c for c in df_sim.columns.tolist() if c != "time"]
        errors = np.random.normal(0, 1, df_sim[cols].shape)
        df_sim[cols] = df_sim[cols] + df_sim[cols] * errors * coef_variation
        df_sim.set_index("time", inplace=True)
        return df_sim

    def simulate_samples(self, parameters: pd.DataFrame, simulation_settings: SimulationSettings):
        """Simulate all rows of ``parameters`` in one launch.  Reference :221-272.

        Returns a dataset with dims (sim, time) and one variable per observable named "[sid]".
        Raises ``RuntimeError`` if any trajectory fails to integrate (roadrunner raises on CVODE
        failure, :242).
        """
        st = simulation_settings
        pb = self.problem
        n = parameters.shape[0]
        pids = [str(c) for c in parameters.columns]

        pb.reset_values()                                     # r.resetAll()            (:229)
        if st.dosage:
            for key, value in st.dosage.items():              # r.setValue(key, value)  (:231-233)
                pb.set_value(key, value)
        pb.set_columns(pids)                                  # r.setValue(pid, row[pid]) (:236-239)
        tgrid = np.linspace(st.start, st.end, int(st.steps) + 1)   # simulate(start, end, steps) (:242-246)
        pb.set_timepoints(tgrid)

        if st.observables:
            obs_ls: list[str] = []
            for observed in st.observables:                   # :250-256 (mutates the ids like the reference)
                if observed.id[0] != "[" or observed.id[-1] != "]":
                    observed.id = "[" + observed.id + "]"
                obs_ls.append(observed.id)
        else:
            obs_ls = [f"[{s}]" for s in pb.state_ids]
        pb.set_observables(obs_ls)

        P = np.ascontiguousarray(parameters.to_numpy(dtype=np.float64).T) if pids else None
        status = np.empty(n, dtype=np.int32)
        steps = np.empty((2, n), dtype=np.int32)
        Y, n_failed = pb.simulate_host(P, n, status=status, steps=steps)   # [sim, time, obs]
        self.last_status, self.last_steps = status, steps
        if n_failed:
            bad = np.nonzero(status)[0]
            raise RuntimeError(
                f"integration failed for {n_failed} of {n} parameter rows (first: row {bad[0]}, "
                f"{runtime.STATUS_TEXT.get(int(status[bad[0]]), status[bad[0]])})")

        data = Y                                                              # [sim, time, obs], as written by the kernel
        noise = st.noise
        if noise is not None and getattr(noise, "add_noise", False):
            # host side, row by row, in the reference's order: every row reseeds the global stream (:260-265)
            for i in range(n):
                seed = np.random.randint(low=0, high=2001)
                np.random.seed(seed)
                errors = np.random.normal(0, 1, data[i].shape)
                data[i] = data[i] + data[i] * errors * noise.cv

        dvars = {name: np.ascontiguousarray(data[:, :, k]) for k, name in enumerate(obs_ls)}
        dset = SimDataset(time=tgrid, data=dvars)
        if _xr is not None:  # pragma: no cover
            return dset.to_xarray()
        return dset
