#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN CODE for the rows of the hot path that are plain numpy / pandas.

`src/parvar/experiments/petab_factory.py` cannot be imported here (it imports roadrunner, xarray, pymetadata ... at module
level, none installed), but three of its functions on the path do not touch any of that:

    lognorm_par_transform          :94-98     
    create_samples_parameters      :101-121   (row a1; includes the loc / scale swap, quirk Q1, and the global seed, Q2)
    ODESampleSimulator.add_noise   :192-219   (row a4; reseeds the global generator per individual, quirks Q2 / Q3)

This script reads the reference file, takes exactly those definitions (and the LognormParameters dataclass, :71-77) out
of its syntax tree, executes them unmodified, and stores their inputs and outputs in tests/golden/reference_sampling_noise.npz.
tests/test_golden.py then holds the oracle's restatement AND the product's host mirror to these vectors bit for bit - these
rows are pinned by the reference itself, not by the oracle.  (The ODE rows a3 / a5-a9 stay unpinned: their arithmetic is
libroadrunner's and AMICI's.)

A second fixture pins the on-disk format either side of the path: the reference's PEtab
writer `create_petab_example` (:320-517) is plain pandas / yaml too.  It is executed the same way on a small data set
(2 groups x 3 individuals x 4 time points x 3 observables; the product's xarray-free `SimDataset` stands in for the
xr.Dataset argument, it offers the four members the writer touches) and the directory it writes is committed as
tests/golden/petab_reference_simple_pk/ (without the SBML file it copies).  tests/test_golden.py checks that the
product's reader parses it and that the product's own writer produces the same tables.

Runs in the build container only (/root/reference does not exist on the GPU box); the fixtures are committed.

usage: python tools/make_golden_reference.py [/root/reference]
"""
import ast
import importlib
import shutil
import sys
import tempfile
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "reference_sampling_noise.npz"


import contextlib
import types


@contextlib.contextmanager
def reference_modules(ref_root: Path, models: Optional[dict] = None, roadrunner_cls=object, dataset_cls=object, concat=None):
    """Execute the reference's `parvar.experiments.{utils, experiment, petab_factory}` from their source files and yield
    the petab_factory module.  What cannot be imported in this container is replaced by inert stand-ins: `roadrunner`
    (`roadrunner_cls`), `xarray` (`dataset_cls`, `concat`), matplotlib, `pymetadata.console`, `rich.progress.track`,
    `pydantic_yaml` (YAML round trip of experiments, unused here); `parvar/__init__.py` (creates result directories at
    import) is bypassed, `parvar.MODELS` is `models`.  sys.modules is restored on exit."""
    saved = dict(sys.modules)
    try:
        def mod(name, **attrs):
            m = types.ModuleType(name)
            m.__path__ = []
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m
        mod("roadrunner", RoadRunner=roadrunner_cls, Integrator=object)
        mod("xarray", Dataset=dataset_cls, concat=concat)
        mod("matplotlib").pyplot = mod("matplotlib.pyplot")
        quiet = types.SimpleNamespace(print=lambda *a, **k: None, rule=lambda *a, **k: None)
        mod("pymetadata").console = mod("pymetadata.console", console=quiet)
        mod("rich").progress = mod("rich.progress", track=lambda it, **k: it)
        py = mod("pydantic_yaml")
        py.parse_yaml_raw_as = py.to_yaml_str = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
        mod("parvar", MODELS=models or {})
        mod("parvar.experiments")
        base = ref_root / "src" / "parvar" / "experiments"
        for name in ("utils", "experiment", "petab_factory"):
            m = types.ModuleType(f"parvar.experiments.{name}")
            m.__file__ = str(base / f"{name}.py")
            sys.modules[m.__name__] = m
            exec(compile((base / f"{name}.py").read_text(), m.__file__, "exec"), m.__dict__)
        pfm = sys.modules["parvar.experiments.petab_factory"]
        pfm.plot_samples = pfm.plot_simulations = lambda *a, **k: None
        yield pfm
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]


def reference_factory(ref_root: Path, model: str) -> dict:
    """`factory_data` of src/parvar/experiments/factories/<model>.py (inside `reference_modules`)."""
    ns: dict = {}
    f = ref_root / "src" / "parvar" / "experiments" / "factories" / f"{model}.py"
    exec(compile(f.read_text(), str(f), "exec"), ns)
    return ns["factory_data"]


def load_reference_functions(ref_root: Path) -> dict:
    src_file = ref_root / "src" / "parvar" / "experiments" / "petab_factory.py"
    tree = ast.parse(src_file.read_text())
    wanted: list[ast.stmt] = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("lognorm_par_transform", "create_samples_parameters"):
            wanted.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "LognormParameters":
            wanted.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "ODESampleSimulator":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "add_noise":
                    sub.decorator_list = []          # a plain function here; the body is untouched
                    wanted.append(sub)
    names = sorted(n.name for n in wanted)
    assert names == ["LognormParameters", "add_noise", "create_samples_parameters", "lognorm_par_transform"], names
    # names the definitions refer to in annotations / at run time; the enums are only dictionary keys and type hints
    ns = {"np": np, "pd": pd, "dataclass": dataclass, "Optional": Optional, "Category": str, "PKPDParameters": str,
          "DistributionType": str}
    module = ast.Module(body=wanted, type_ignores=[])
    exec(compile(module, str(src_file), "exec"), ns)
    return ns


# sampling definitions of the three reference factories: (group, parameter) -> (loc, scale) of `pars_true`
FACTORIES = {
    "simple_chain": {"MALE": {"k1": (1.0, 1)}, "FEMALE": {"k1": (2.0, 1)}},                               # factories/simple_chain.py:12-25
    "simple_pk": {"MALE": {"CL": (1.0, 1), "k_abs": (0.5, 1)}, "FEMALE": {"CL": (0.5, 1), "k_abs": (1.0, 1)}},   # simple_pk.py:19-44, order :100-121
    "icg_body_flat": {"MALE": {"BW": (75.0, 10), "LI__ICGIM_Vmax": (0.0369598840327503, 0.01)},             # icg_body_flat.py:23-49
                      "FEMALE": {"BW": (65.0, 10), "LI__ICGIM_Vmax": (0.02947, 0.01)}},
}


def oracle_trajectory(oracle, model: str, overrides: dict, t: np.ndarray) -> np.ndarray:
    """What stands in for the ODE solve when the reference's workflow is executed (and when the tests replay it)."""
    if model == "simple_pk":
        return oracle.simple_pk_exact(overrides, t)
    return oracle.simulate(model, overrides, t)


def petab_fixture_data():
    """The small data set of the PEtab fixture: deterministic numbers, no simulator involved."""
    t = np.array([0.0, 10.0, 50.0, 100.0])
    species = ["y_cent", "y_gut", "y_peri"]
    data = {}
    for gi, grp in enumerate(("MALE", "FEMALE")):
        y = np.empty((3, len(t), len(species)))
        for i in range(3):
            for k in range(len(species)):
                y[i, :, k] = (0.1 + 0.05 * k + 0.01 * i + 0.2 * gi) * np.exp(-(0.01 + 0.005 * k) * t) * (1 + 0.03 * ((i + k) % 3))
        data[grp] = y
    priors = {"k_abs_MALE": (0.5, 1.0), "k_abs_FEMALE": (1.0, 1.0), "CL_MALE": (1.0, 1.0), "CL_FEMALE": (0.5, 1.0)}
    return t, species, data, priors


def write_petab_with_reference_code(ref_root: Path, out_dir: Path) -> None:
    sys.path.insert(0, str(ROOT))
    ds = importlib.import_module("parameter-variability_b200.experiments.dataset")
    pkg = importlib.import_module("parameter-variability_b200")
    src_file = ref_root / "src" / "parvar" / "experiments" / "petab_factory.py"
    tree = ast.parse(src_file.read_text())
    wanted = [n for n in tree.body
              if (isinstance(n, ast.FunctionDef) and n.name == "create_petab_example")
              or (isinstance(n, ast.ClassDef) and n.name == "Category")
              or (isinstance(n, ast.Assign) and any(isinstance(tg, ast.Name) and tg.id in ("MEASUREMENT_UNIT_COLUMN", "MEASUREMENT_TIME_UNIT_COLUMN")
                                                    for tg in n.targets))]
    assert len(wanted) == 4, [getattr(n, "name", "assign") for n in wanted]
    import enum
    import typing
    import yaml
    ns = {"np": np, "pd": pd, "yaml": yaml, "shutil": shutil, "Path": Path, "Enum": enum.Enum, "Optional": Optional,
          "Union": typing.Union, "xr": type("xr", (), {"Dataset": object}), "Group": object}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), str(src_file), "exec"), ns)
    Category = ns["Category"]
    t, species, data, priors = petab_fixture_data()

    class Estimation:                      # the two members of parvar.experiments.experiment.Group the writer uses (:439-444)
        parameters = True

    class GroupStub:
        estimation = Estimation()

        def __init__(self, gid):
            self.gid = gid

        def get_parameter(self, kind, par, key):
            assert kind == "estimation"
            return priors[f"{par}_{self.gid}"][0 if key == "loc" else 1]

    dfs = {Category[g]: ds.SimDataset(t, {f"[{sp}]": y[:, :, k] for k, sp in enumerate(species)}) for g, y in data.items()}
    with tempfile.TemporaryDirectory() as tmp:
        xp = Path(tmp) / "refuid"
        ns["create_petab_example"](dfs, [GroupStub("MALE"), GroupStub("FEMALE")], xp / "petab", ["k_abs", "CL"],
                                   Path(pkg.MODELS["simple_pk"]))
        (xp / "petab" / Path(pkg.MODELS["simple_pk"]).name).unlink()        # the copied model is not part of the fixture
        if out_dir.exists():
            shutil.rmtree(out_dir)
        shutil.copytree(xp, out_dir)
    print(f"wrote {out_dir} with the reference's create_petab_example")


def study_definitions(ref_root: Path, out_file: Path) -> None:
    """The sweep grid (factories/__init__.py:12-20) and the three factories' experiment definitions - sampling and prior
    distributions, sample counts, steps, noise, observables, parameter order - as the reference's own modules build them.
    experiment.py (pydantic models) and factories/*.py are executed unmodified (`reference_modules`)."""
    import json
    out = {}
    with reference_modules(ref_root):
        ns: dict = {}
        f = ref_root / "src" / "parvar" / "experiments" / "factories" / "__init__.py"
        exec(compile(f.read_text(), str(f), "exec"), ns)
        out["definitions_all"] = ns["definitions_all"]["all"]
        out["definitions_minimal"] = ns["definitions_minimal"]
        for model in ("simple_chain", "simple_pk", "icg_body_flat"):
            fd = reference_factory(ref_root, model)
            dump = lambda pars: {k: v.model_dump(mode="json") for k, v in pars.items()}
            out[model] = {"exp_base": fd["exp_base"].model_dump(mode="json"), "pars_true": dump(fd["pars_true"]),
                          "pars_biased": {k: dump(v) for k, v in fd["pars_biased"].items()}}
    # the estimation settings: the constant keywords of `settings = dict(...)` in optimize_experiment (petab_optimization.py:294-306)
    tree = ast.parse((ref_root / "src" / "parvar" / "optimization" / "petab_optimization.py").read_text())
    for node in ast.walk(tree):
        if (isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id == "settings"
                and isinstance(node.value, ast.Call) and getattr(node.value.func, "id", "") == "dict"):
            out["optimization_settings"] = {kw.arg: ast.literal_eval(kw.value) for kw in node.value.keywords
                                            if isinstance(kw.value, ast.Constant)}
    assert out["optimization_settings"]["n_starts"] == 100
    out_file.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print(f"wrote {out_file} from the reference's experiment.py and factories/")


def workflow_fixture(ref_root: Path, out_file: Path, model: str = "simple_pk", n_samples: int = 4) -> None:
    """The reference's whole data-generation workflow - `create_petab_for_experiment` (:548-621) with the real
    `ODESampleSimulator.simulate_samples` (:221-272), `create_samples_parameters`, `add_noise`, `create_petab_example` - run
    from the reference's own module source for a small experiment per model (its factory's `exp_base` with a few individuals
    per group, 5 time points, cv 0.1).  Only the ODE solve is not the reference's: `roadrunner.RoadRunner` is a stand-in that
    answers `simulate()` with the oracle (exact matrix exponential for simple_pk, scipy at rtol 1e-12 for the other two -
    `oracle_trajectory` below, which the tests call as well); xarray is a stand-in that builds the
    product's `SimDataset`; plotting is switched off.  What the fixture pins is everything AROUND the solve: which random
    numbers are drawn in which order (one global legacy stream: seed 1234, log-normal draws per group and parameter, then per
    individual randint -> reseed -> normal), dose / parameter hand-over to the simulator, bracketed observable ids, the
    start / end / steps -> time grid convention, and the table assembly."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import parvar_oracle as oracle
    ds = importlib.import_module("parameter-variability_b200.experiments.dataset")
    pkg = importlib.import_module("parameter-variability_b200")
    captured: dict = {}

    class FakeIntegrator:
        def setSetting(self, key, value):
            captured.setdefault("integrator", {})[key] = value

    class SimResult(np.ndarray):            # roadrunner's NamedArray: a 2-D array with .colnames
        colnames: list

    class FakeRoadRunner:
        def __init__(self, path):
            assert Path(path).stem == model
            self.integrator = FakeIntegrator()
            self.values: dict = {}
            self.calls: list = []

        def getIds(self):
            return list(oracle.MODELS[model].states)

        def resetAll(self):
            self.values = {}

        def setValue(self, key, value):
            self.values[key] = float(value)
            captured.setdefault("set_order", []).append(key)

        def simulate(self, start, end, steps):
            t = np.linspace(start, end, steps + 1)
            y = oracle_trajectory(oracle, model, dict(self.values), t)
            res = np.column_stack([t, y]).view(SimResult)
            res.colnames = ["time"] + [f"[{sid}]" for sid in oracle.MODELS[model].states]
            return res

    class Dataset(ds.SimDataset):
        def to_netcdf(self, path):
            captured.setdefault("dsets", {})[Path(path).stem] = self

    def concat(frames, dim):
        return Dataset(frames[0].index.to_numpy(), {c: np.stack([f[c].to_numpy() for f in frames]) for c in frames[0].columns},
                       sim=np.asarray(dim))

    saved_to_xarray = pd.DataFrame.to_xarray
    try:
        pd.DataFrame.to_xarray = lambda self: self           # consumed by the xarray stand-in's concat only
        with reference_modules(ref_root, models={model: Path(pkg.MODELS[model])}, roadrunner_cls=FakeRoadRunner,
                               dataset_cls=Dataset, concat=concat) as pfm:
            xp = reference_factory(ref_root, model)["exp_base"].model_copy(deep=True)
            xp.id = "wf"
            for g in xp.groups:
                g.sampling.n_samples, g.sampling.steps, g.sampling.noise.cv = n_samples, 4, 0.1
            with tempfile.TemporaryDirectory() as tmp:
                yaml_file = pfm.create_petab_for_experiment(xp, Path(tmp), show_plot=False)
                meas = pd.read_csv(Path(yaml_file).parent / "petab" / f"measurements_{model}.tsv", sep="\t",
                                   float_precision="round_trip")
                cond = pd.read_csv(Path(yaml_file).parent / "petab" / f"conditions_{model}.tsv", sep="\t")
    finally:
        pd.DataFrame.to_xarray = saved_to_xarray
    keys = list(captured["dsets"])
    out = {"time": captured["dsets"][keys[0]]["time"].values, "groups": np.array(keys),
           "abs_tol": captured["integrator"]["absolute_tolerance"], "rel_tol": captured["integrator"]["relative_tolerance"]}
    for key, dset in captured["dsets"].items():
        out[f"vars/{key}"] = np.array(list(dset.data_vars))
        out[f"noisy/{key}"] = np.stack([dset.data_vars[v] for v in dset.data_vars], axis=2)       # [sim, time, observable]
    n_set = len(captured["set_order"]) // (len(keys) * n_samples)            # setValue calls per simulated row, in order (:231-239)
    out["set_order_first_row"] = np.array(captured["set_order"][:n_set])
    out["condition_columns"] = np.array(list(cond.columns))
    out["measurement"] = meas["measurement"].to_numpy()
    out["measurement_condition"] = meas["simulationConditionId"].to_numpy().astype(str)
    out["measurement_observable"] = meas["observableId"].to_numpy().astype(str)
    out["measurement_time"] = meas["time"].to_numpy()
    np.savez_compressed(out_file, **out)
    print(f"wrote {out_file} from the reference's create_petab_for_experiment (groups {keys})")

    # ---- drop-in check of boundary B1 (not a fixture): the SAME reference workflow once more, now with the reference's
    # ODESampleSimulator class replaced by the product's mirror (its real simulate_samples; the kernel wrapper answered by the
    # oracle like RoadRunner above).  The reference's code must run unchanged and produce the same data.
    pf = importlib.import_module("parameter-variability_b200.experiments.petab_factory")

    class FakeProblem:
        state_ids = list(oracle.MODELS[model].states)

        def reset_values(self):
            self.over = {}

        def set_value(self, key, value):
            self.over[key] = value

        def set_columns(self, ids):
            self.cols = list(ids)

        def set_timepoints(self, t):
            self.t = np.asarray(t, dtype=float)

        def set_observables(self, obs):
            self.idx = [self.state_ids.index(o.strip("[]")) for o in obs]

        def simulate_host(self, P, n_rows, status=None, steps=None):
            Y = np.stack([oracle_trajectory(oracle, model, {**self.over, **{c: P[j, i] for j, c in enumerate(self.cols)}}, self.t)[:, self.idx]
                          for i in range(n_rows)])
            status[:] = 0
            steps[:] = 0
            return Y, 0

    second: dict = {}

    class MirrorSimulator(pf.ODESampleSimulator):
        def __init__(self, model_path, abs_tol=1e-6, rel_tol=1e-6):          # the reference's call: ODESampleSimulator(model_path=...)
            self.model_name, self.problem = model, FakeProblem()

        def simulate_samples(self, parameters, simulation_settings):
            dset = super().simulate_samples(parameters, simulation_settings)
            second[len(second)] = dset          # the reference goes on to call dset.to_netcdf(...): SimDataset's own (scipy, NetCDF-3)
            return dset

    with reference_modules(ref_root, models={model: Path(pkg.MODELS[model])}) as pfm:
        pfm.ODESampleSimulator = MirrorSimulator
        xp = reference_factory(ref_root, model)["exp_base"].model_copy(deep=True)
        xp.id = "wf"
        for g in xp.groups:
            g.sampling.n_samples, g.sampling.steps, g.sampling.noise.cv = n_samples, 4, 0.1
        with tempfile.TemporaryDirectory() as tmp:
            yaml_file = pfm.create_petab_for_experiment(xp, Path(tmp), show_plot=False)
            meas2 = pd.read_csv(Path(yaml_file).parent / "petab" / f"measurements_{model}.tsv", sep="\t", float_precision="round_trip")
    assert len(second) == len(keys)
    for k, key in enumerate(keys):
        assert list(second[k].data_vars) == list(captured["dsets"][key].data_vars)
        for v in second[k].data_vars:
            assert np.array_equal(second[k].data_vars[v], captured["dsets"][key].data_vars[v]), (key, v)
    assert meas2.equals(meas)
    print(f"  drop-in check: the reference workflow with the product's ODESampleSimulator class gives the same {model} data")


PIDS = ["k_abs_MALE", "CL_FEMALE", "LI__ICGIM_Vmax_MALE", "BW_OLD", "k1", "k1_", "_MALE", "a_b_c", "par_gr\u00fcn", "x__y", "",
        "k_abs_MALE\n", "UPPER_lower_123"]


def pid_helpers(ref_root: Path, out_file: Path) -> None:
    """experiments/utils.py (importable as is): how a PEtab parameter id splits into (parameter, group) in the results table
    (`get_parameter_from_pid`, `get_group_from_pid`, used at petab_optimization.py:254-262)."""
    import json
    ns: dict = {"__name__": "ref_utils"}
    f = ref_root / "src" / "parvar" / "experiments" / "utils.py"
    exec(compile(f.read_text(), str(f), "exec"), ns)
    out = {"pids": PIDS, "group": [ns["get_group_from_pid"](p) for p in PIDS],
           "parameter": [ns["get_parameter_from_pid"](p) for p in PIDS],
           "uuid_lengths": {str(n): len(ns["uuid_alphanumeric"](n)) for n in (1, 20, 45)}}
    out_file.write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {out_file} from the reference's experiments/utils.py")


def analysis_inputs():
    """Deterministic result frames in the product's results schema (driver.results_df) for three small experiments."""
    xps = [{"id": f"uid{k}", "model": "simple_pk", "prior_type": ["exact_prior", "prior_biased", "no_prior"][k], "samples": 20,
            "timepoints": 10 + k, "noise_cv": 0.05,
            "sampling": {"MALE": {"CL": (1.0, 1.0), "k_abs": (0.5, 1.0)}, "FEMALE": {"CL": (0.5, 1.0), "k_abs": (1.0, 1.0)}}} for k in range(3)]
    frames = {}
    for k, xp in enumerate(xps):
        rows = []
        for j, pid in enumerate(("k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE")):
            grp = pid.rsplit("_", 1)[1]
            rows.append({"id": xp["id"], "group": grp, "parameter": pid[: -len(grp) - 1], "pid": pid, "mean": 1.0 + 0.1 * j + k,
                         "std": 0.1 + 0.01 * j, "median": 0.9 + 0.1 * j + k, "ess": 100.0 + j, "n_samples": 50,
                         "optim_duration": 0.1 * (k + 1), "sampler_duration": 0.2, "values": (np.arange(8.0).reshape(2, 4) + j) / 7.0,
                         "hdi_low": 0.8 + k, "hdi_high": 1.2 + k})
        frames[xp["id"]] = pd.DataFrame(rows)
    return xps, frames


def analysis_fixture(ref_root: Path, out_dir: Path) -> None:
    """The other end of the path: the results files.  The PRODUCT writes `definitions.tsv` + `<uid>_results.tsv`
    (optimization/driver.py::write_reference_layout), and the reference's own `join_optimization_results`
    (src/parvar/analysis/utils.py:28-98, executed from its source, console stubbed) reads and joins them.  The files that
    went in and the table that came out are committed; the test checks that the product still writes exactly these files."""
    import types
    sys.path.insert(0, str(ROOT))
    drv = importlib.import_module("parameter-variability_b200.optimization.driver")
    saved = dict(sys.modules)
    try:
        pm, pmc = types.ModuleType("pymetadata"), types.ModuleType("pymetadata.console")
        pmc.console = types.SimpleNamespace(print=lambda *a, **k: None, rule=lambda *a, **k: None)
        sys.modules["pymetadata"], sys.modules["pymetadata.console"] = pm, pmc
        ns: dict = {"__name__": "ref_analysis_utils"}
        f = ref_root / "src" / "parvar" / "analysis" / "utils.py"
        exec(compile(f.read_text(), str(f), "exec"), ns)
        xps, frames = analysis_inputs()
        with tempfile.TemporaryDirectory() as tmp:
            base = drv.write_reference_layout(Path(tmp), "all", xps, frames)
            joined = ns["join_optimization_results"](Path(tmp), "all")
            assert len(joined) == 2 * 4, len(joined)          # the no_prior experiment is dropped by the reference (:91)
            (base / "definitions_results.parquet").unlink()
            if out_dir.exists():
                shutil.rmtree(out_dir)
            shutil.copytree(base, out_dir)
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
    print(f"wrote {out_dir}: product-written results joined by the reference's join_optimization_results")


def experiment_factory_fixture(ref_root: Path, out_file: Path) -> None:
    """All 3 x 2160 experiments of the study as the reference's own `model_experiment_factory` (:623-700) builds them from
    `definitions_all` and each factory's data: per experiment the prior type, samples, steps (= timepoints - 1), noise cv and
    the prior (loc, scale) that `create_petab_example` would write for every (group, parameter) - looked up the reference's
    way, `group.get_parameter("estimation", par, ...)` = first match in the group's estimation list (which for MALE also holds
    the FEMALE entries: `g.id in par` is a substring test, :686-693)."""
    out: dict = {}
    with reference_modules(ref_root) as pfm:
        ns: dict = {}
        f = ref_root / "src" / "parvar" / "experiments" / "factories" / "__init__.py"
        exec(compile(f.read_text(), str(f), "exec"), ns)
        definition = ns["definitions_all"]["all"]
        for model in ("simple_chain", "simple_pk", "icg_body_flat"):
            xps = pfm.model_experiment_factory(**definition, **reference_factory(ref_root, model)).experiments
            params = [p.id for p in xps[0].groups[0].get_parameter_list("sampling")]          # create_petab_for_experiment :611
            out[f"{model}/params"] = np.array(params)
            out[f"{model}/groups"] = np.array([g.id for g in xps[0].groups])
            out[f"{model}/prior_type"] = np.array([x.prior_type for x in xps])
            out[f"{model}/samples"] = np.array([[g.sampling.n_samples for g in x.groups] for x in xps])
            out[f"{model}/steps"] = np.array([[g.sampling.steps for g in x.groups] for x in xps])
            out[f"{model}/cv"] = np.array([[g.sampling.noise.cv for g in x.groups] for x in xps])
            out[f"{model}/prior"] = np.array([[[[g.get_parameter("estimation", p, k) for k in ("loc", "scale")] for p in params]
                                               for g in x.groups] for x in xps])                # [experiment, group, parameter, 2]
            out[f"{model}/n_estimation_entries"] = np.array([[len(g.estimation.parameters) for g in x.groups] for x in xps])
    np.savez_compressed(out_file, **out)
    print(f"wrote {out_file}: {len(out['simple_pk/prior_type'])} experiments per model from the reference's model_experiment_factory")


def definitions_table_fixture(ref_root: Path, out_file: Path) -> None:
    """`definitions.tsv` as the reference writes it: `create_petabs_for_definitions` (:702-716) -> `create_petabs` (:520-546)
    -> `PETabExperimentList.to_dataframe` (experiment.py:170-204), run on `definitions_minimal` for simple_pk (3 prior types x
    {5, 10} time points; what the reference's own tests/experiments/test_factories.py runs) with the stand-ins of
    `workflow_fixture`.  Only the table is kept."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import parvar_oracle as oracle
    ds = importlib.import_module("parameter-variability_b200.experiments.dataset")
    pkg = importlib.import_module("parameter-variability_b200")

    class SimResult(np.ndarray):
        colnames: list

    class Integrator:
        def setSetting(self, key, value):
            pass

    class RoadRunner:
        def __init__(self, path):
            self.integrator, self.values = Integrator(), {}

        def getIds(self):
            return list(oracle.SIMPLE_PK.states)

        def resetAll(self):
            self.values = {}

        def setValue(self, key, value):
            self.values[key] = float(value)

        def simulate(self, start, end, steps):
            t = np.linspace(start, end, steps + 1)
            res = np.column_stack([t, oracle.simple_pk_exact(dict(self.values), t)]).view(SimResult)
            res.colnames = ["time"] + [f"[{sid}]" for sid in oracle.SIMPLE_PK.states]
            return res

    class Dataset(ds.SimDataset):
        def to_netcdf(self, path):
            pass

    def concat(frames, dim):
        return Dataset(frames[0].index.to_numpy(), {c: np.stack([f[c].to_numpy() for f in frames]) for c in frames[0].columns},
                       sim=np.asarray(dim))

    saved_to_xarray = pd.DataFrame.to_xarray
    try:
        pd.DataFrame.to_xarray = lambda self: self
        with reference_modules(ref_root, models={"simple_pk": Path(pkg.MODELS["simple_pk"])}, roadrunner_cls=RoadRunner,
                               dataset_cls=Dataset, concat=concat) as pfm:
            ns: dict = {}
            f = ref_root / "src" / "parvar" / "experiments" / "factories" / "__init__.py"
            exec(compile(f.read_text(), str(f), "exec"), ns)
            counter = iter(range(10 ** 6))
            pfm.uuid_alphanumeric = lambda length=20: f"xp{next(counter):0{length - 2}d}"     # reproducible ids (uuid4 in the reference)
            with tempfile.TemporaryDirectory() as tmp:
                pfm.create_petabs_for_definitions(ns["definitions_minimal"], Path(tmp), reference_factory(ref_root, "simple_pk"))
                table = Path(tmp) / "xps" / "timepoints" / pfm.DEFINITIONS_FILENAME
                out_file.write_bytes(table.read_bytes())
                n_dirs = len([d for d in table.parent.iterdir() if d.is_dir()])
    finally:
        pd.DataFrame.to_xarray = saved_to_xarray
    print(f"wrote {out_file}: definitions.tsv of {n_dirs} experiments written by the reference's create_petabs")


def main():
    ref_root = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    definitions_table_fixture(ref_root, ROOT / "tests" / "golden" / "reference_definitions_minimal_simple_pk.tsv")
    experiment_factory_fixture(ref_root, ROOT / "tests" / "golden" / "reference_experiment_factory.npz")
    analysis_fixture(ref_root, ROOT / "tests" / "golden" / "analysis_reference_join")
    pid_helpers(ref_root, ROOT / "tests" / "golden" / "reference_pid_helpers.json")
    workflow_fixture(ref_root, ROOT / "tests" / "golden" / "reference_workflow_simple_pk.npz")
    workflow_fixture(ref_root, ROOT / "tests" / "golden" / "reference_workflow_simple_chain.npz", model="simple_chain", n_samples=3)
    workflow_fixture(ref_root, ROOT / "tests" / "golden" / "reference_workflow_icg_body_flat.npz", model="icg_body_flat", n_samples=3)
    study_definitions(ref_root, ROOT / "tests" / "golden" / "reference_study_definitions.json")
    write_petab_with_reference_code(ref_root, ROOT / "tests" / "golden" / "petab_reference_simple_pk")
    ref = load_reference_functions(ref_root)
    L = ref["LognormParameters"]
    out: dict[str, np.ndarray] = {}

    # ---- lognorm_par_transform -------------------------------------------------------------------------------------------
    pairs = np.array([(1.0, 1.0), (0.5, 1.0), (1.0, 0.5), (75.0, 10.0), (10.0, 75.0), (0.0369598840327503, 0.01), (2.0, 3.0)])
    out["lpt_in"] = pairs
    out["lpt_out"] = np.array([ref["lognorm_par_transform"](a, b) for a, b in pairs])

    # ---- create_samples_parameters: every factory at n = 7 and n = 160 (the largest sweep value), default seed; and one
    # call with another seed.  LognormParameters(mu=loc, sigma=scale, n=...) as the reference builds them (:567-571) ----------
    for model, groups in FACTORIES.items():
        for n in (7, 160):
            pars = {g: {p: L(n=n, sigma=sc, mu=lo) for p, (lo, sc) in d.items()} for g, d in groups.items()}
            s = ref["create_samples_parameters"](pars)
            for g, d in s.items():
                for p, v in d.items():
                    out[f"samples/{model}/{n}/{g}/{p}"] = np.asarray(v)
    pars = {g: {p: L(n=5, sigma=sc, mu=lo) for p, (lo, sc) in d.items()} for g, d in FACTORIES["simple_pk"].items()}
    s = ref["create_samples_parameters"](pars, seed=99)
    out["samples_seed99/simple_pk/MALE/CL"] = np.asarray(s["MALE"]["CL"])

    # ---- add_noise: a small trajectory table, explicit seed; then the reference's per-individual pattern - the global stream
    # left behind by create_samples_parameters decides the seeds (seed=None -> np.random.randint(0, 2001)) --------------------
    t = np.linspace(0.0, 100.0, 6)
    table = pd.DataFrame({"time": t, "[y_cent]": np.exp(-0.1 * t) * 0.3, "[y_gut]": np.exp(-t), "[y_peri]": 0.1 * t * np.exp(-0.05 * t)})
    out["noise_in"] = table.to_numpy()
    noisy = ref["add_noise"](table.set_index("time"), 0.1, "normal", seed=17)
    out["noise_seed17_cv0.1"] = noisy.reset_index().to_numpy()
    pars = {g: {p: L(n=3, sigma=sc, mu=lo) for p, (lo, sc) in d.items()} for g, d in FACTORIES["simple_pk"].items()}
    ref["create_samples_parameters"](pars)                       # seeds the global generator (1234) and draws
    seq = []
    for _ in range(4):                                           # four individuals, cv = 0.05 (factories/simple_pk.py:108)
        seq.append(ref["add_noise"](table.set_index("time"), 0.05, "normal").reset_index().to_numpy())
    out["noise_sequence_cv0.05"] = np.stack(seq)

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({len(out)} arrays) from {ref_root}")


if __name__ == "__main__":
    main()
