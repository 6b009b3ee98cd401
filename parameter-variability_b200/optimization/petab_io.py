"""Reader for the PEtab v1 problems the reference writes.

``create_petab_example`` (reference src/parvar/experiments/petab_factory.py:320-517) writes, per experiment,
``petab.yaml`` next to a ``petab/`` directory holding ``measurements_<model>.tsv``, ``conditions_<model>.tsv``,
``parameters_<model>.tsv``, ``observables_<model>.tsv`` and a copy of the SBML file.  ``load_petab_problem``
turns such a directory into a ``PetabObjective`` so that the GPU objective can be driven from the reference's
existing experiment folders; the reference does this with ``petab.v1.Problem.from_yaml`` + pyPESTO's
``PetabImporter`` (src/parvar/optimization/petab_optimization.py:41-48).  Only the table features the reference
emits are supported (pandas + yaml, no petab/libsbml dependency):

  measurements : observableId, simulationConditionId, measurement, time            (:404-417)
  conditions   : conditionId, <par> -> "<par>_<cond>", <species> -> initial value   (:350-376)
  parameters   : parameterId, parameterScale=lin, lowerBound, upperBound, estimate,
                 objectivePriorType=parameterScaleNormal, objectivePriorParameters="loc;scale"   (:428-451)
  observables  : observableId, observableFormula=<species>, noiseFormula (number), noiseDistribution  (:388-398)
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd
import yaml

from .. import runtime
from .petab_optimization import GroupData, PetabObjective


@dataclass
class PetabTables:
    yaml_path: Path
    model_path: Path
    measurements: pd.DataFrame
    conditions: pd.DataFrame
    parameters: pd.DataFrame
    observables: pd.DataFrame


def read_petab_tables(yaml_path: str | Path) -> PetabTables:
    yaml_path = Path(yaml_path)
    with open(yaml_path) as f:
        cfg = yaml.safe_load(f)
    if cfg.get("format_version") not in (1, "1"):
        raise ValueError(f"{yaml_path}: only PEtab format_version 1 is supported")
    base = yaml_path.parent
    prob = cfg["problems"][0]

    def tsv(rel):
        # round_trip: pandas' default float parser is off by a few ulp; the reference writes repr() floats, read them exactly
        return pd.read_csv(base / rel, sep="\t", float_precision="round_trip")

    return PetabTables(
        yaml_path=yaml_path,
        model_path=base / prob["sbml_files"][0],
        measurements=pd.concat([tsv(p) for p in prob["measurement_files"]]),
        conditions=pd.concat([tsv(p) for p in prob["condition_files"]]),
        parameters=tsv(cfg["parameter_file"]),
        observables=pd.concat([tsv(p) for p in prob["observable_files"]]),
    )


@dataclass
class ParsedPetab:
    """What the tables of one reference experiment directory say (no GPU object involved)."""
    model: str
    parameters: list            # model parameters mapped to estimated "<par>_<cond>" ids, condition-table column order
    groups: list                # GroupData per condition
    timepoints: np.ndarray
    species: list               # observableFormula per observable = the observed species
    sigma: np.ndarray           # noiseFormula per observable
    priors: dict                # x_name -> (loc, scale)
    literal_values: dict        # species columns of the condition table (quirk Q4: observed species -> 0.0, no dose)
    x_names: list
    lb: np.ndarray
    ub: np.ndarray
    x_nominal: np.ndarray


def parse_petab_problem(yaml_path: str | Path) -> ParsedPetab:
    """Read and check the tables ``create_petab_example`` writes (petab_factory.py:320-517)."""
    from ..experiments.petab_factory import resolve_model
    tb = read_petab_tables(yaml_path)
    model = resolve_model(tb.model_path)

    obs = tb.observables
    observable_ids = list(obs["observableId"])
    species = [str(f) for f in obs["observableFormula"]]
    if any(str(d) != "normal" for d in obs.get("noiseDistribution", pd.Series(["normal"] * len(obs)))):
        raise NotImplementedError("only noiseDistribution = normal (what the reference writes) is read from PEtab files")
    sigma = np.array([float(x) for x in obs["noiseFormula"]])
    if "observableTransformation" in obs and any(str(t) != "lin" for t in obs["observableTransformation"]):
        raise NotImplementedError("only observableTransformation = lin is supported")

    par = tb.parameters
    if any(str(s) != "lin" for s in par["parameterScale"]):
        raise NotImplementedError("only parameterScale = lin (what the reference writes for estimated parameters) is supported")
    est = par[par["estimate"] == 1]
    cond = tb.conditions
    cond_ids = list(cond["conditionId"])
    # condition columns mapping a model parameter to an estimated parameter id "<par>_<cond>"
    model_pars = [c for c in cond.columns if c not in ("conditionId", "conditionName")
                  and all(str(v) in set(est["parameterId"]) for v in cond[c])]
    x_names = [f"{p}_{c}" for p in model_pars for c in cond_ids]
    for p in model_pars:
        for c, v in zip(cond_ids, cond[p]):
            if str(v) != f"{p}_{c}":
                raise NotImplementedError(f"condition table maps {p} to {v!r}; expected '{p}_{c}' as the reference writes")

    meas = tb.measurements
    times = np.sort(meas["time"].unique())
    groups = []
    for c in cond_ids:
        m = meas[meas["simulationConditionId"] == c]
        per_obs = []
        n_ind = None
        for oid in observable_ids:
            mo = m[m["observableId"] == oid]
            # individuals are written one after another, each with every time point once (:378-417)
            counts = mo.groupby("time").size()
            k = int(counts.max()) if len(counts) else 0
            arr = np.full((k, len(times)), np.nan)
            seen = {t: 0 for t in times}
            for t, y in zip(mo["time"].to_numpy(), mo["measurement"].to_numpy()):
                arr[seen[t], np.searchsorted(times, t)] = y
                seen[t] += 1
            per_obs.append(arr)
            n_ind = k if n_ind is None else max(n_ind, k)
        y = np.full((n_ind, len(times), len(observable_ids)), np.nan)
        for k_obs, arr in enumerate(per_obs):
            y[:arr.shape[0], :, k_obs] = arr
        groups.append(GroupData(c, y))

    priors = {}
    if "objectivePriorType" in est.columns:
        for pid, ptype, pp in zip(est["parameterId"], est["objectivePriorType"], est["objectivePriorParameters"]):
            if isinstance(ptype, str) and ptype:
                if ptype not in ("parameterScaleNormal", "normal"):
                    raise NotImplementedError(f"prior type {ptype}")
                loc, scale = (float(v) for v in str(pp).split(";"))
                priors[pid] = (loc, scale)

    # the species columns of the condition table (identical for every condition in the reference's files)
    literal_values = {}
    for col in cond.columns:
        if col in ("conditionId", "conditionName") or col in model_pars:
            continue
        vals = set(float(v) for v in cond[col])
        if len(vals) != 1:
            raise NotImplementedError(f"condition-specific initial value for {col}")
        literal_values[col] = vals.pop()
    lb = {pid: float(v) for pid, v in zip(par["parameterId"], par["lowerBound"])}
    ub = {pid: float(v) for pid, v in zip(par["parameterId"], par["upperBound"])}
    nominal = par.set_index("parameterId")["nominalValue"]
    return ParsedPetab(model=model, parameters=model_pars, groups=groups, timepoints=times, species=species, sigma=sigma,
                       priors=priors, literal_values=literal_values, x_names=x_names,
                       lb=np.array([lb[n] for n in x_names]), ub=np.array([ub[n] for n in x_names]),
                       x_nominal=np.array([float(nominal.loc[n]) for n in x_names]))


def load_petab_problem(yaml_path: str | Path, mode: str = "literal", dosage: Optional[dict[str, float]] = None,
                       device: int = 0, **objective_kwargs) -> PetabObjective:
    """Build the GPU objective for one reference experiment directory.

    ``mode="literal"`` applies the condition table as written - which sets every observed species' initial
    value to 0.0 and carries no dose (quirk Q4) - exactly what AMICI would simulate for these files;
    ``mode="intended"`` ignores those species overrides and applies ``dosage`` instead.
    """
    if mode not in ("literal", "intended"):
        raise ValueError("mode must be 'literal' or 'intended'")
    pp = parse_petab_problem(yaml_path)
    objective = PetabObjective(pp.model, pp.parameters, pp.groups, pp.timepoints, pp.species, sigma=pp.sigma, priors=pp.priors,
                               mode="intended", dosage=dosage if mode == "intended" else None, device=device, **objective_kwargs)
    if mode == "literal":
        for col, v in pp.literal_values.items():
            objective.problem.set_value(col, v)
    assert objective.x_names == pp.x_names
    objective.lb, objective.ub, objective.x_nominal = pp.lb, pp.ub, pp.x_nominal
    return objective


def write_petab_like_reference(directory: Path, model_path: Path, groups: list[GroupData], timepoints, species: list[str],
                               parameters: list[str], priors: dict[str, tuple[float, float]]) -> Path:
    """Write the four tables + yaml in the layout and column set of ``create_petab_example`` (:320-517).
    The product's PEtab writer (``experiments/workflow.py`` uses it; byte-identical to the reference's five files,
    tests/test_golden.py); kept next to the reader so the two are maintained together."""
    directory = Path(directory)
    petab = directory / "petab"
    petab.mkdir(parents=True, exist_ok=True)
    stem = Path(model_path).stem
    rows = []
    for g in groups:
        for i in range(g.y.shape[0]):
            for k, sp in enumerate(species):
                for j, t in enumerate(timepoints):
                    rows.append({"observableId": f"{sp}_observable", "preequilibrationConditionId": None,
                                 "simulationConditionId": g.id, "measurement": g.y[i, j, k], "measurementUnit": "mmole/l",
                                 "time": t, "timeUnit": "second", "observableParameters": None, "noiseParameters": None})
    pd.DataFrame(rows).to_csv(petab / f"measurements_{stem}.tsv", sep="\t", index=False)
    cond = []
    for g in groups:
        c = {"conditionId": g.id, "conditionName": ""}
        for p in parameters:
            c[p] = f"{p}_{g.id}"
        for sp in species:
            c[sp] = 0.0                                     # quirk Q4 (:369-376)
        cond.append(c)
    pd.DataFrame(cond).to_csv(petab / f"conditions_{stem}.tsv", sep="\t", index=False)
    pars = []
    for p in parameters:
        for g in groups:
            pid = f"{p}_{g.id}"
            row = {"parameterId": pid, "parameterName": pid, "parameterScale": "lin", "lowerBound": 0.0, "upperBound": 1e8,
                   "nominalValue": 1, "estimate": 1, "parameterUnit": "l/min"}
            if pid in priors:
                row["objectivePriorType"] = "parameterScaleNormal"
                row["objectivePriorParameters"] = f"{priors[pid][0]};{priors[pid][1]}"
            pars.append(row)
    pd.DataFrame(pars).to_csv(petab / f"parameters_{stem}.tsv", sep="\t", index=False)
    obs = [{"observableId": f"{sp}_observable", "observableFormula": sp, "observableName": sp, "noiseDistribution": "normal",
            "noiseFormula": 1, "observableTransformation": "lin", "observableUnit": "mmol/l"} for sp in species]
    pd.DataFrame(obs).to_csv(petab / f"observables_{stem}.tsv", sep="\t", index=False)
    import shutil
    shutil.copy(model_path, petab / Path(model_path).name)
    cfg = {"format_version": 1, "parameter_file": f"petab/parameters_{stem}.tsv",
           "problems": [{"condition_files": [f"petab/conditions_{stem}.tsv"], "measurement_files": [f"petab/measurements_{stem}.tsv"],
                         "observable_files": [f"petab/observables_{stem}.tsv"], "sbml_files": [f"petab/{Path(model_path).name}"]}]}
    with open(directory / "petab.yaml", "w") as f:
        yaml.dump(cfg, f, default_flow_style=False)
    return directory / "petab.yaml"
