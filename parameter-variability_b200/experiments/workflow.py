"""``create_petab_for_experiment`` on the GPU path (reference src/parvar/experiments/petab_factory.py:548-621).

The reference's function draws the parameter samples of every group (:576), simulates them group by group with
``ODESampleSimulator`` (:582-599), stores the datasets and writes the PEtab problem (:610-619).  This is the same
sequence on the mirrors of this package - sampling on the host with numpy's global legacy stream, ONE batched launch
per group, host-side noise in the reference's order, PEtab tables in the reference's layout - without the pydantic
experiment objects (out of scope): the experiment is passed as plain values, or as any object
with the reference's attribute structure (``PETabExperiment``), which ``experiment_to_spec`` flattens.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

from .. import MODELS
from ..optimization.petab_io import write_petab_like_reference
from ..optimization.petab_optimization import GroupData
from .experiment import Noise, Observable
from .petab_factory import LognormParameters, ODESampleSimulator, SimulationSettings, create_samples_parameters


@dataclass
class GroupSpec:
    id: str                                        # "MALE", "FEMALE", ...
    sampling: dict[str, tuple[float, float]]       # parameter -> (loc, scale) of the sampling distribution, in order
    estimation: dict[str, tuple[float, float]] = field(default_factory=dict)   # parameter -> prior (loc, scale)
    n_samples: int = 10                            # experiment.py:74-77 defaults
    steps: int = 20
    tend: float = 100.0
    cv: float = 0.0
    add_noise: bool = True


@dataclass
class ExperimentSpec:
    id: str
    model: str
    groups: list[GroupSpec]
    observables: list[str]
    dosage: Optional[dict[str, float]] = None


def experiment_to_spec(xp) -> ExperimentSpec:
    """Flatten a reference ``PETabExperiment`` (experiment.py:106-156) - duck typed, pydantic not required."""
    groups = []
    for g in xp.groups:
        samp = {p.id: (p.distribution.parameters["loc"], p.distribution.parameters["scale"]) for p in g.sampling.parameters}
        est = {p.id: (p.distribution.parameters["loc"], p.distribution.parameters["scale"]) for p in g.estimation.parameters}
        groups.append(GroupSpec(id=g.id, sampling=samp, estimation=est, n_samples=g.sampling.n_samples, steps=g.sampling.steps,
                                tend=g.sampling.tend, cv=g.sampling.noise.cv, add_noise=g.sampling.noise.add_noise))
    obs = [o.id.strip("[]") for o in xp.groups[0].sampling.observables]
    return ExperimentSpec(id=xp.id, model=xp.model, groups=groups, observables=obs, dosage=xp.dosage)


def create_petab_for_experiment(experiment: ExperimentSpec, directory: Path, device: int = 0, seed: Optional[int] = 1234):
    """-> (petab.yaml path, {group: dataset}, {group: parameter DataFrame}).  Reference :548-621."""
    xp_path = Path(directory) / f"{experiment.id}"
    xp_path.mkdir(parents=True, exist_ok=True)
    sbml_path = MODELS[experiment.model]                                                       # :558
    samples_dsn = {g.id: {par: LognormParameters(mu=loc, sigma=scale, n=g.n_samples)              # :562-574
                          for par, (loc, scale) in g.sampling.items()} for g in experiment.groups}
    samples = create_samples_parameters(samples_dsn, seed=seed)                                  # :576
    simulator = ODESampleSimulator(model_path=sbml_path, device=device)                          # :582
    dsets, frames = {}, {}
    for g in experiment.groups:                                                                  # :584-599
        settings = SimulationSettings(start=0.0, end=g.tend, steps=g.steps, dosage=experiment.dosage,
                                      noise=Noise(add_noise=g.add_noise, cv=g.cv),
                                      observables=[Observable(id=o) for o in experiment.observables])
        parameters = pd.DataFrame({par: v for par, v in samples[g.id].items()})
        dsets[g.id] = simulator.simulate_samples(parameters, simulation_settings=settings)
        frames[g.id] = parameters
    # PEtab problem (:610-619): measurements per group in dataset order, priors from the estimation block
    groups_data = []
    for g in experiment.groups:
        ds = dsets[g.id]
        y = np.stack([ds[f"[{o}]"].values for o in experiment.observables], axis=2)             # [sim, time, obs]
        groups_data.append(GroupData(g.id, y))
    g0 = experiment.groups[0]
    tgrid = np.linspace(0.0, g0.tend, g0.steps + 1)
    params = list(g0.sampling.keys())                                                            # :611
    priors = {f"{par}_{g.id}": ls for g in experiment.groups for par, ls in g.estimation.items()}
    yaml_file = write_petab_like_reference(xp_path, sbml_path, groups_data, tgrid, experiment.observables, params, priors)
    return yaml_file, dsets, frames
