"""Committed golden vectors (tests/golden/, made by tools/make_golden.py from the oracle):
  - not-gpu: the oracle (Python and C restatements) still reproduces them  -> pins the oracle against drift;
  - gpu:     the CUDA path reproduces them through the C ABI."""
from pathlib import Path

import numpy as np
import pytest

from conftest import parity_violation, pkg

G = Path(__file__).resolve().parent / "golden"


def test_oracle_reproduces_golden(oracle):
    g = np.load(G / "simple_pk_c2.npz")
    for b in (0, 17, 63):
        over = {"k_abs": g["k_abs"][b], "CL": g["CL"][b]}
        np.testing.assert_allclose(oracle.simple_pk_exact(over, g["t"]), g["Y"][:, :, b], rtol=1e-12, atol=1e-300)
        assert oracle.loglik_normal(g["Y"][:, :, b], g["data"], float(g["sigma"])) == pytest.approx(g["logp"][b], rel=1e-13)
    g = np.load(G / "icg_body_flat_c4.npz")
    Y = oracle.simulate("icg_body_flat", {"IVDOSE_icg": 10.0, "BW": g["BW"][5], "LI__ICGIM_Vmax": g["Vmax"][5]}, g["t"])
    assert parity_violation(Y, g["Y"][:, :, 5], rtol=1e-8, atol=1e-12) <= 1.0
    g = np.load(G / "sampling_simple_pk.npz")
    s = oracle.create_samples_parameters({"MALE": {"CL": (20, 1.0, 1.0), "k_abs": (20, 1.0, 0.5)},
                                          "FEMALE": {"CL": (20, 1.0, 0.5), "k_abs": (20, 1.0, 1.0)}})
    np.testing.assert_array_equal(s["FEMALE"]["k_abs"], g["k_abs_FEMALE"])
    np.testing.assert_array_equal(oracle.add_noise(g["y"], 0.05), g["noisy"])


def test_host_mirror_sampling_matches_golden():
    """create_samples_parameters of the host mirror = the reference's stream, bit for bit."""
    pf = pkg("experiments.petab_factory")
    g = np.load(G / "sampling_simple_pk.npz")
    L = pf.LognormParameters
    s = pf.create_samples_parameters({
        pf.Category.MALE: {pf.PKPDParameters.CL: L(n=20, sigma=1.0, mu=1.0), pf.PKPDParameters.k_abs: L(n=20, sigma=1.0, mu=0.5)},
        pf.Category.FEMALE: {pf.PKPDParameters.CL: L(n=20, sigma=1.0, mu=0.5), pf.PKPDParameters.k_abs: L(n=20, sigma=1.0, mu=1.0)}})
    np.testing.assert_array_equal(s[pf.Category.MALE][pf.PKPDParameters.CL], g["CL_MALE"])
    np.testing.assert_array_equal(s[pf.Category.FEMALE][pf.PKPDParameters.k_abs], g["k_abs_FEMALE"])


def test_c_oracle_matches_golden():
    """The C restatement (the bench's CPU baseline) integrates the same systems: variable-order BDF at tight
    tolerance vs the golden trajectories, and at the reference's 1e-6/1e-6 within the accuracy those settings give."""
    import oracle_c
    g = np.load(G / "simple_pk_c2.npz")
    obs = ["y_gut", "y_cent", "y_peri"]
    Y, lp, nf, st = oracle_c.simulate_batch("simple_pk", {"k_abs": g["k_abs"], "CL": g["CL"]}, g["t"], obs,
                                            rtol=1e-9, atol=1e-12, data=g["data"], sigma=float(g["sigma"]))
    assert nf == 0 and parity_violation(Y, g["Y"]) <= 1.0
    np.testing.assert_allclose(lp, g["logp"], rtol=1e-6)
    Y6, _, nf, _ = oracle_c.simulate_batch("simple_pk", {"k_abs": g["k_abs"], "CL": g["CL"]}, g["t"], obs, rtol=1e-6, atol=1e-6)
    assert nf == 0 and np.max(np.abs(Y6 - g["Y"])) < 2e-5          # quirk Q9: the reference's own data accuracy
    g = np.load(G / "icg_body_flat_c4.npz")
    states = list(g["states"])
    Y, _, nf, _ = oracle_c.simulate_batch("icg_body_flat", {"BW": g["BW"], "LI__ICGIM_Vmax": g["Vmax"]}, g["t"], states,
                                          overrides={"IVDOSE_icg": 10.0}, rtol=1e-9, atol=1e-13)
    assert nf == 0 and parity_violation(Y, g["Y"]) <= 1.0
    g = np.load(G / "simple_chain.npz")
    Y, _, nf, _ = oracle_c.simulate_batch("simple_chain", {"k1": g["k1"]}, g["t"], ["S1", "S2"], rtol=1e-9, atol=1e-12)
    assert nf == 0 and parity_violation(Y, g["Y"]) <= 1.0


def test_c_oracle_rhs_matches_python_oracle(oracle):
    import ctypes
    import oracle_c
    lib = oracle_c.load()
    rs = np.random.RandomState(0)
    for name in ("simple_chain", "simple_pk", "icg_body_flat"):
        m, states, params, defaults = oracle_c.model_info(name)
        om = oracle.MODELS[name]
        assert states == om.states
        p = np.array(defaults) * rs.lognormal(0, 0.2, len(defaults))
        y = rs.rand(len(states)) * 1e-2
        f = np.zeros(len(states))
        dp = ctypes.POINTER(ctypes.c_double)
        lib.pvo_rhs(m, y.ctypes.data_as(dp), p.ctypes.data_as(dp), f.ctypes.data_as(dp))
        ref = np.array(om.rhs(0.0, y, dict(zip(params, p))), dtype=float)
        np.testing.assert_allclose(f, ref, rtol=1e-12, atol=1e-300)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cuda_reproduces_golden_simple_pk(cuda_lib, rt):
    g = np.load(G / "simple_pk_c2.npz")
    pb = rt.Problem("simple_pk")
    pb.set_solver(rtol=1e-7, atol=1e-10)
    pb.set_columns(["k_abs", "CL"])
    pb.set_timepoints(g["t"])
    pb.set_observables(list(g["states"]))
    th = np.stack([g["k_abs"], g["CL"]])
    data = g["data"]
    pb.set_data(np.stack([np.ones_like(data), data, data * data])[..., None], float(g["sigma"]))
    lp_out = np.empty(len(g["k_abs"]))
    Y, nf = pb.simulate_host(th, th.shape[1], logp=lp_out)
    assert nf == 0 and parity_violation(Y, np.moveaxis(g["Y"], -1, 0)) <= 1.0       # golden files are [T, obs, B]
    np.testing.assert_allclose(lp_out, g["logp"], rtol=1e-6)
    pb.set_solver(rtol=1e-9, atol=1e-13)
    lp, gr, _, nf = pb.logp_host(th, grad=True)
    np.testing.assert_allclose(lp, g["logp"], rtol=1e-6)
    np.testing.assert_allclose(gr, g["grad"], rtol=1e-6, atol=1e-6 * np.abs(g["grad"]).max())


@pytest.mark.gpu
def test_cuda_reproduces_golden_chain_and_icg(cuda_lib, rt):
    import torch
    g = np.load(G / "simple_chain.npz")
    pb = rt.Problem("simple_chain")
    pb.set_columns(["k1"])
    pb.set_timepoints(g["t"])
    pb.set_observables(["S1", "S2"])
    Y, nf = pb.simulate_host(g["k1"][None, :], len(g["k1"]))
    assert nf == 0 and parity_violation(Y, np.moveaxis(g["Y"], -1, 0)) <= 1.0
    B, T = len(g["k1"]), len(g["t"])
    pb.set_solver(rtol=1e-9, atol=1e-12)
    P = torch.from_numpy(g["k1"][None, :]).cuda()
    Yd = torch.empty((B, T, 2), dtype=torch.float64, device="cuda")
    S = torch.empty((B, T, 2, 1), dtype=torch.float64, device="cuda")
    pb.simulate_sens(P, Yd, S)
    torch.cuda.synchronize()
    assert parity_violation(S[..., 0].cpu().numpy(), np.moveaxis(g["S"][:, :, 0, :], -1, 0)) <= 1.0

    g = np.load(G / "icg_body_flat_c4.npz")
    pb = rt.Problem("icg_body_flat")
    pb.set_value("IVDOSE_icg", 10.0)
    pb.set_columns(["BW", "LI__ICGIM_Vmax"])
    pb.set_timepoints(g["t"])
    pb.set_observables(list(g["states"]))
    th = np.stack([g["BW"], g["Vmax"]])
    Y, nf = pb.simulate_host(th, th.shape[1])
    assert nf == 0 and parity_violation(Y, np.moveaxis(g["Y"], -1, 0)) <= 1.0
    pb.set_observables(list(g["obs"]))
    pb.set_solver(rtol=1e-9, atol=1e-14)
    data = g["data"]
    pb.set_data(np.stack([np.ones_like(data), data, data * data])[..., None], float(g["sigma"]))
    lp, gr, _, nf = pb.logp_host(th, grad=True)
    assert nf == 0
    np.testing.assert_allclose(lp, g["logp"], rtol=1e-6)
    np.testing.assert_allclose(gr[:, :4], g["grad"], rtol=1e-5)


# ---- vectors produced by the reference's own code (tools/make_golden_reference.py: the pure numpy / pandas functions of
# src/parvar/experiments/petab_factory.py executed out of its syntax tree) ---------------------------------------------
REF_FACTORIES = {
    "simple_chain": {"MALE": {"k1": (1.0, 1)}, "FEMALE": {"k1": (2.0, 1)}},
    "simple_pk": {"MALE": {"CL": (1.0, 1), "k_abs": (0.5, 1)}, "FEMALE": {"CL": (0.5, 1), "k_abs": (1.0, 1)}},
    "icg_body_flat": {"MALE": {"BW": (75.0, 10), "LI__ICGIM_Vmax": (0.0369598840327503, 0.01)},
                      "FEMALE": {"BW": (65.0, 10), "LI__ICGIM_Vmax": (0.02947, 0.01)}},
}


def test_oracle_sampling_and_noise_match_the_reference_code(oracle):
    """Rows a1 / a4 of SURVEY.md section 8a are pinned by the reference itself: lognorm_par_transform (:94-98),
    create_samples_parameters (:101-121, with the loc / scale swap and the global seed) and add_noise (:192-219, with the
    per-individual reseeding) of the oracle reproduce the outputs of the reference's functions bit for bit."""
    g = np.load(G / "reference_sampling_noise.npz")
    for (a, b), want in zip(g["lpt_in"], g["lpt_out"]):
        np.testing.assert_array_equal(np.array(oracle.lognorm_par_transform(a, b)), want)
    for model, groups in REF_FACTORIES.items():
        for n in (7, 160):
            s = oracle.create_samples_parameters({grp: {p: (n, sc, lo) for p, (lo, sc) in d.items()} for grp, d in groups.items()})
            for grp, d in groups.items():
                for p in d:
                    np.testing.assert_array_equal(s[grp][p], g[f"samples/{model}/{n}/{grp}/{p}"])
    s = oracle.create_samples_parameters({grp: {p: (5, sc, lo) for p, (lo, sc) in d.items()}
                                          for grp, d in REF_FACTORIES["simple_pk"].items()}, seed=99)
    np.testing.assert_array_equal(s["MALE"]["CL"], g["samples_seed99/simple_pk/MALE/CL"])
    table = g["noise_in"]
    np.testing.assert_array_equal(oracle.add_noise(table[:, 1:], 0.1, seed=17), g["noise_seed17_cv0.1"][:, 1:])
    oracle.create_samples_parameters({grp: {p: (3, sc, lo) for p, (lo, sc) in d.items()}
                                      for grp, d in REF_FACTORIES["simple_pk"].items()})
    for k in range(4):
        np.testing.assert_array_equal(oracle.add_noise(table[:, 1:], 0.05), g["noise_sequence_cv0.05"][k][:, 1:])


def test_host_mirror_sampling_and_noise_match_the_reference_code():
    """The product's host mirror (experiments/petab_factory.py) against the same reference-generated vectors."""
    import pandas as pd
    pf = pkg("experiments.petab_factory")
    g = np.load(G / "reference_sampling_noise.npz")
    L = pf.LognormParameters
    for (a, b), want in zip(g["lpt_in"], g["lpt_out"]):
        np.testing.assert_array_equal(np.array(pf.lognorm_par_transform(a, b)), want)
    for model, groups in REF_FACTORIES.items():
        for n in (7, 160):
            s = pf.create_samples_parameters({grp: {p: L(n=n, sigma=sc, mu=lo) for p, (lo, sc) in d.items()}
                                              for grp, d in groups.items()})
            for grp, d in groups.items():
                for p in d:
                    np.testing.assert_array_equal(s[grp][p], g[f"samples/{model}/{n}/{grp}/{p}"])
    table = g["noise_in"]
    cols = ["time", "[y_cent]", "[y_gut]", "[y_peri]"]
    df = pd.DataFrame(table, columns=cols).set_index("time")
    noisy = pf.ODESampleSimulator.add_noise(df, 0.1, seed=17)
    np.testing.assert_array_equal(noisy.reset_index().to_numpy(), g["noise_seed17_cv0.1"])
    pf.create_samples_parameters({grp: {p: L(n=3, sigma=sc, mu=lo) for p, (lo, sc) in d.items()}
                                  for grp, d in REF_FACTORIES["simple_pk"].items()})
    for k in range(4):
        noisy = pf.ODESampleSimulator.add_noise(pd.DataFrame(table, columns=cols).set_index("time"), 0.05)
        np.testing.assert_array_equal(noisy.reset_index().to_numpy(), g["noise_sequence_cv0.05"][k])


def _petab_fixture_data():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_reference", G.parent.parent / "tools" / "make_golden_reference.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.petab_fixture_data()


def test_petab_reader_and_writer_against_reference_written_directory(tmp_path, monkeypatch):
    """tests/golden/petab_reference_simple_pk/ was written by the reference's own create_petab_example
    (petab_factory.py:320-517, executed by tools/make_golden_reference.py with the product's SimDataset as its data set).
    The product's reader must turn it back into exactly the data it was written from - parameters, groups, individuals,
    time grid, observables, priors, bounds, and the literal-mode species overrides (quirk Q4) - and the product's own
    writer must produce the same files byte for byte."""
    import shutil
    io = pkg("optimization.petab_io")
    P = pkg()
    t, species, data, priors = _petab_fixture_data()
    xp = tmp_path / "refuid"
    shutil.copytree(G / "petab_reference_simple_pk", xp)
    shutil.copy(P.MODELS["simple_pk"], xp / "petab" / "simple_pk.xml")      # the writer copies the model next to the tables
    seen = {}

    class ObjectiveStub:                      # PetabObjective's constructor needs a GPU; the reader's output is its arguments
        def __init__(self, model, pars, groups, times, spec_, **kw):
            seen.update(model=model, pars=list(pars), groups=groups, times=np.asarray(times), species=list(spec_), kw=kw, set=[])
            self.x_names = [f"{p}_{g.id}" for p in pars for g in groups]
            self.problem = type("PB", (), {"set_value": lambda _s, k, v: seen["set"].append((k, v))})()

    monkeypatch.setattr(io, "PetabObjective", ObjectiveStub)
    obj = io.load_petab_problem(xp / "petab.yaml", mode="literal")
    assert seen["model"] == "simple_pk" and seen["pars"] == ["k_abs", "CL"] and seen["species"] == species
    np.testing.assert_array_equal(seen["times"], t)
    assert [g.id for g in seen["groups"]] == ["MALE", "FEMALE"]
    for g in seen["groups"]:
        np.testing.assert_array_equal(g.y, data[g.id])                      # [individual, time, observable], bit for bit
    assert seen["kw"]["priors"] == priors
    np.testing.assert_array_equal(seen["kw"]["sigma"], [1.0, 1.0, 1.0])     # noiseFormula = 1 (quirk Q5)
    assert sorted(seen["set"]) == [("y_cent", 0.0), ("y_gut", 0.0), ("y_peri", 0.0)]     # quirk Q4, literal mode
    assert obj.x_names == ["k_abs_MALE", "k_abs_FEMALE", "CL_MALE", "CL_FEMALE"]
    np.testing.assert_array_equal(obj.lb, np.zeros(4))
    np.testing.assert_array_equal(obj.ub, np.full(4, 1e8))
    np.testing.assert_array_equal(obj.x_nominal, np.ones(4))
    # the same tables parsed without any GPU object (what the batched driver uses), and handed on to the batch objective
    pp = io.parse_petab_problem(xp / "petab.yaml")
    assert (pp.model, pp.parameters, pp.species, pp.x_names) == ("simple_pk", ["k_abs", "CL"], species, obj.x_names)
    assert pp.priors == priors and pp.literal_values == {"y_cent": 0.0, "y_gut": 0.0, "y_peri": 0.0}
    np.testing.assert_array_equal(pp.sigma, [1.0, 1.0, 1.0])
    drv = pkg("optimization.driver")
    got = {}
    monkeypatch.setattr(drv, "PetabObjectiveBatch", lambda *a, **kw: got.update(a=a, kw=kw) or type("B", (), {})())
    drv._batch_from_yaml([xp / "petab.yaml", xp / "petab.yaml"], "literal", None, 0)
    assert got["a"][0] == "simple_pk" and len(got["a"][2]) == 2 and got["kw"]["initial_values"] == pp.literal_values
    np.testing.assert_array_equal(got["kw"]["sigma"], pp.sigma)
    assert got["kw"]["priors"] == [priors, priors]

    mine = tmp_path / "mine"
    groups = [io.GroupData("MALE", data["MALE"]), io.GroupData("FEMALE", data["FEMALE"])]
    io.write_petab_like_reference(mine, P.MODELS["simple_pk"], groups, t, species, ["k_abs", "CL"], priors)
    for rel in ("petab.yaml", "petab/measurements_simple_pk.tsv", "petab/conditions_simple_pk.tsv",
                "petab/parameters_simple_pk.tsv", "petab/observables_simple_pk.tsv"):
        assert (mine / rel).read_bytes() == (G / "petab_reference_simple_pk" / rel).read_bytes(), rel


def test_study_definitions_match_the_reference_factories():
    """tests/golden/reference_study_definitions.json is a dump of the objects the reference's own experiment.py and
    factories/*.py build (tools/make_golden_reference.py executes those modules).  The study driver's transcription of them
    (tools/run_study.py: sweep grid, sampling and prior distributions, parameter order, observables, dose) and the sampling
    definitions used by the fixtures above must say the same."""
    import importlib.util
    import json
    ref = json.loads((G / "reference_study_definitions.json").read_text())
    spec = importlib.util.spec_from_file_location("run_study", G.parent.parent / "tools" / "run_study.py")
    rs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rs)
    assert rs.DEFINITIONS_ALL == ref["definitions_all"]
    assert len(ref["definitions_all"]["prior_types"]) * len(ref["definitions_all"]["samples"]) * \
        len(ref["definitions_all"]["timepoints"]) * len(ref["definitions_all"]["noise_cvs"]) == 2160

    def ls(p):
        return (p["distribution"]["parameters"]["loc"], p["distribution"]["parameters"]["scale"])

    for model, mine in rs.MODELS.items():
        r = ref[model]
        xb = r["exp_base"]
        assert xb["model"] == model
        assert (xb["dosage"] or None) == mine["dosage"]
        assert [g["id"] for g in xb["groups"]] == list(mine["sampling"])
        for g in xb["groups"]:
            smp = g["sampling"]
            assert list(mine["sampling"][g["id"]].items()) == [(p["id"], ls(p)) for p in smp["parameters"]]     # draw order
            assert list(REF_FACTORIES[model][g["id"]].items()) == [(p["id"], ls(p)) for p in smp["parameters"]]
            assert [o["id"] for o in smp["observables"]] == mine["observables"]
            assert smp["tend"] == 100.0 and smp["noise"]["type"] == "normal"                                    # quirks Q3, Q8
            assert [p["id"] for p in smp["parameters"]] == mine["params"]       # PEtab parameter order = sampling order (:611)
            for p in mine["params"]:
                key = f"{p}_{g['id']}" if f"{p}_{g['id']}" in r["pars_true"] else g["id"]     # simple_chain keys by group only
                assert mine["priors"]["exact_prior"][g["id"]][p] == ls(r["pars_true"][key])
                for kind in ("prior_biased", "prior_incorrect"):
                    assert mine["priors"][kind][g["id"]][p] == ls(r["pars_biased"][kind][key])


def _golden_tools():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_reference", G.parent.parent / "tools" / "make_golden_reference.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


WORKFLOW_CASES = [("simple_pk", 4), ("simple_chain", 3), ("icg_body_flat", 3)]


@pytest.mark.parametrize("model,n", WORKFLOW_CASES)
def test_oracle_flow_reproduces_the_reference_workflow(oracle, model, n):
    """tests/golden/reference_workflow_<model>.npz: the reference's own create_petab_for_experiment (:548-621), with its real
    simulate_samples / create_samples_parameters / add_noise / create_petab_example, run by tools/make_golden_reference.py for
    a 2-group, n-individual, 5-time-point experiment of every model; only the ODE solve inside RoadRunner.simulate was answered
    by the oracle.  The oracle's restatement of that flow - the checker of
    tests/test_gpu_estimation.py::test_create_petab_for_experiment_like_reference - must reproduce it bit for bit: same draws
    from the one global stream in the same order, dose before parameters, same grid, observable order and table order."""
    import json
    traj = _golden_tools().oracle_trajectory
    g = np.load(G / f"reference_workflow_{model}.npz")
    xb = json.loads((G / "reference_study_definitions.json").read_text())[model]["exp_base"]
    assert float(g["abs_tol"]) == 1e-6 and float(g["rel_tol"]) == 1e-6             # quirk Q9: the reference's data tolerances
    t = np.linspace(0.0, 100.0, 5)                                                 # start 0, end tend = 100, steps 4 (quirk Q8)
    np.testing.assert_array_equal(g["time"], t)
    groups = REF_FACTORIES[model]
    assert [str(k) for k in g["groups"]] == [f"Category.{k}" for k in groups]
    dosage = xb["dosage"] or {}
    assert [str(k) for k in g["set_order_first_row"]] == list(dosage) + list(groups["MALE"])      # setValue order (:231-239)
    observed = [o["id"] for o in xb["groups"][0]["sampling"]["observables"]]
    assert [str(c) for c in g["condition_columns"]] == ["conditionId", "conditionName"] + list(groups["MALE"]) + observed   # Q4: no dose
    samples = oracle.create_samples_parameters({grp: {p: (n, sc, lo) for p, (lo, sc) in d.items()} for grp, d in groups.items()})
    states = oracle.MODELS[model].states
    rows = []
    for grp in groups:
        names = [str(v) for v in g[f"vars/Category.{grp}"]]
        assert names == [f"[{o}]" for o in observed]                               # bracketed ids, factory order (:250-258)
        idx = [states.index(o) for o in observed]
        want = g[f"noisy/Category.{grp}"]
        for i in range(n):
            clean = traj(oracle, model, {**dosage, **{p: samples[grp][p][i] for p in groups[grp]}}, t)[:, idx]
            noisy = oracle.add_noise(clean, 0.1)                                   # randint -> reseed -> normal, per individual
            np.testing.assert_array_equal(noisy, want[i])
            for k, o in enumerate(observed):
                rows += [(grp, o + "_observable", tj, noisy[j, k]) for j, tj in enumerate(t)]
    # the measurement table: group, individual, observable, time (petab_factory.py:378-417)
    assert [r[0] for r in rows] == [str(c) for c in g["measurement_condition"]]
    assert [r[1] for r in rows] == [str(c) for c in g["measurement_observable"]]
    np.testing.assert_array_equal([r[2] for r in rows], g["measurement_time"])
    np.testing.assert_array_equal([r[3] for r in rows], g["measurement"])


@pytest.mark.parametrize("model,n", WORKFLOW_CASES)
def test_product_workflow_host_logic_reproduces_the_reference_workflow(oracle, tmp_path, monkeypatch, model, n):
    """The product's data-generation workflow (experiments/workflow.py::create_petab_for_experiment with the mirror's real
    ODESampleSimulator.simulate_samples: sampling, hand-over of dose / columns / grid / observables to the kernel wrapper,
    per-row host noise, PEtab tables) against the reference's own workflow (reference_workflow_<model>.npz).  Only the
    kernel call is replaced - `Problem.simulate_host` answers with the oracle, as RoadRunner.simulate did when the fixture
    was made - so the two must agree bit for bit, measurement table included."""
    import json
    import pandas as pd
    traj = _golden_tools().oracle_trajectory
    wf = pkg("experiments.workflow")
    pf = pkg("experiments.petab_factory")
    g = np.load(G / f"reference_workflow_{model}.npz")
    xb = json.loads((G / "reference_study_definitions.json").read_text())[model]["exp_base"]

    class FakeProblem:                                       # the members of runtime.Problem simulate_samples touches
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
            Y = np.stack([traj(oracle, model, {**self.over, **{c: P[j, i] for j, c in enumerate(self.cols)}}, self.t)[:, self.idx]
                          for i in range(n_rows)])
            status[:] = 0
            steps[:] = 0
            return Y, 0

    def make_simulator(model_path, device=0):
        sim = pf.ODESampleSimulator.__new__(pf.ODESampleSimulator)      # no GPU: skip the constructor, keep the methods
        sim.model_name, sim.problem = model, FakeProblem()
        return sim

    monkeypatch.setattr(wf, "ODESampleSimulator", make_simulator)

    def ls(p):
        return (p["distribution"]["parameters"]["loc"], p["distribution"]["parameters"]["scale"])

    groups = [wf.GroupSpec(id=grp["id"], sampling={p["id"]: ls(p) for p in grp["sampling"]["parameters"]},
                           estimation={p["id"]: ls(p) for p in grp["estimation"]["parameters"]},
                           n_samples=n, steps=4, tend=grp["sampling"]["tend"], cv=0.1) for grp in xb["groups"]]
    xp = wf.ExperimentSpec(id="wf", model=model, groups=groups,
                           observables=[o["id"] for o in xb["groups"][0]["sampling"]["observables"]], dosage=xb["dosage"])
    yaml_file, dsets, frames = wf.create_petab_for_experiment(xp, tmp_path)
    for grp in groups:
        names = [str(v) for v in g[f"vars/Category.{grp.id}"]]
        assert list(dsets[grp.id].data_vars) == names
        got = np.stack([dsets[grp.id][v].values for v in names], axis=2)
        np.testing.assert_array_equal(got, g[f"noisy/Category.{grp.id}"])
        np.testing.assert_array_equal(dsets[grp.id]["time"].values, g["time"])
    meas = pd.read_csv(tmp_path / "wf" / "petab" / f"measurements_{model}.tsv", sep="\t", float_precision="round_trip")
    np.testing.assert_array_equal(meas["measurement"].to_numpy(), g["measurement"])
    np.testing.assert_array_equal(meas["time"].to_numpy(), g["measurement_time"])
    assert list(meas["simulationConditionId"]) == [str(c) for c in g["measurement_condition"]]
    assert list(meas["observableId"]) == [str(c) for c in g["measurement_observable"]]
    cond = pd.read_csv(tmp_path / "wf" / "petab" / f"conditions_{model}.tsv", sep="\t")
    assert list(cond.columns) == [str(c) for c in g["condition_columns"]]


def test_pid_helpers_match_the_reference_utils():
    """get_group_from_pid / get_parameter_from_pid of the results writer (optimization/driver.py) against the outputs of the
    reference's experiments/utils.py functions (tests/golden/reference_pid_helpers.json), edge cases included."""
    import json
    drv = pkg("optimization.driver")
    ref = json.loads((G / "reference_pid_helpers.json").read_text())
    for pid, grp, par in zip(ref["pids"], ref["group"], ref["parameter"]):
        assert drv.get_group_from_pid(pid) == grp, pid
        assert drv.get_parameter_from_pid(pid) == par, pid


def test_results_files_are_what_the_reference_analysis_joined(tmp_path):
    """tests/golden/analysis_reference_join/: results files written by the product (driver.write_reference_layout) and the
    table the reference's own join_optimization_results (analysis/utils.py:28-98) made of them
    (tools/make_golden_reference.py).  The product must still write exactly those files, and the joined table carries the
    sampling distribution and the posterior summaries where the reference's plots look for them."""
    import importlib.util
    import pandas as pd
    spec = importlib.util.spec_from_file_location("make_golden_reference", G.parent.parent / "tools" / "make_golden_reference.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    drv = pkg("optimization.driver")
    xps, frames = mod.analysis_inputs()
    base = drv.write_reference_layout(tmp_path, "all", xps, frames)
    ref = G / "analysis_reference_join"
    for rel in ["definitions.tsv"] + [f"optimization_results/uid{k}_results.tsv" for k in range(3)]:
        assert (base / rel).read_bytes() == (ref / rel).read_bytes(), rel
    joined = pd.read_csv(ref / "definitions_results.tsv", sep="\t")
    assert len(joined) == 8 and set(joined["prior_type"]) == {"exact_prior", "prior_biased"}       # no_prior rows are dropped (:91)
    for _, row in joined.iterrows():
        xp = next(x for x in xps if x["id"] == row["id"])
        loc, scale = xp["sampling"][row["group"]][row["parameter"]]
        assert (row["sample_loc"], row["sample_scale"]) == (loc, scale)
        src = frames[row["id"]]
        src = src[(src["group"] == row["group"]) & (src["parameter"] == row["parameter"])].iloc[0]
        assert row["bayes_sampler_mean"] == src["mean"] and row["bayes_sampler_median"] == src["median"]
        assert row["ess"] == src["ess"] and row["hdi_low"] == src["hdi_low"] and row["hdi_high"] == src["hdi_high"]


def test_study_enumeration_matches_the_reference_experiment_factory():
    """tests/golden/reference_experiment_factory.npz: the 3 x 2160 experiments the reference's own model_experiment_factory
    (petab_factory.py:623-700) builds from definitions_all.  The study driver (tools/run_study.py) must enumerate the same
    experiments in the same order - prior type, samples, steps = timepoints - 1 (quirk Q8), noise cv - with the same PEtab
    parameter order and the same prior (loc, scale) per (group, parameter)."""
    import importlib.util
    import itertools
    g = np.load(G / "reference_experiment_factory.npz")
    spec = importlib.util.spec_from_file_location("run_study", G.parent.parent / "tools" / "run_study.py")
    rs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rs)
    D = rs.DEFINITIONS_ALL
    defs = list(itertools.product(D["prior_types"], D["samples"], D["timepoints"], D["noise_cvs"]))     # run_study.py:main
    assert len(defs) == 2160
    for model, mine in rs.MODELS.items():
        groups = [str(x) for x in g[f"{model}/groups"]]
        params = [str(x) for x in g[f"{model}/params"]]
        assert groups == list(mine["sampling"]) and params == mine["params"]
        assert [d[0] for d in defs] == [str(x) for x in g[f"{model}/prior_type"]]
        for k in range(len(groups)):
            np.testing.assert_array_equal([d[1] for d in defs], g[f"{model}/samples"][:, k])
            np.testing.assert_array_equal([d[2] - 1 for d in defs], g[f"{model}/steps"][:, k])
            np.testing.assert_array_equal([d[3] for d in defs], g[f"{model}/cv"][:, k])
        prior = np.array([[[mine["priors"][d[0]][grp][p] for p in params] for grp in groups] for d in defs], dtype=float)
        np.testing.assert_array_equal(prior, g[f"{model}/prior"])


def test_estimation_settings_match_the_reference():
    """optimize_experiment's settings (petab_optimization.py:294-306, read out of the reference's syntax tree): multistart and
    sampler sizes, tolerances, seed.  Plot creation and the engine name are the two the GPU driver sets differently on purpose."""
    import json
    ref = json.loads((G / "reference_study_definitions.json").read_text())["optimization_settings"]
    mine = pkg("optimization.driver").DEFAULT_SETTINGS
    assert set(ref) - {"create_optimization_plots"} <= set(mine)
    for key, value in ref.items():
        if key != "create_optimization_plots":
            assert mine[key] == value, key
    assert (ref["n_starts"], ref["n_samples"], ref["n_chains"]) == (100, 5000, 4)


def test_definitions_table_matches_the_reference_writer():
    """tests/golden/reference_definitions_minimal_simple_pk.tsv was written by the reference's own
    create_petabs_for_definitions -> create_petabs -> PETabExperimentList.to_dataframe for `definitions_minimal` (the scenario
    of the reference's tests/experiments/test_factories.py).  The product's definitions_dataframe must write the same bytes."""
    import io
    import itertools
    import json
    drv = pkg("optimization.driver")
    ref_bytes = (G / "reference_definitions_minimal_simple_pk.tsv").read_bytes()
    minimal = json.loads((G / "reference_study_definitions.json").read_text())["definitions_minimal"]["timepoints"]
    xps = [{"id": f"xp{k:018d}", "model": "simple_pk", "prior_type": prior, "samples": 20, "timepoints": n_t, "noise_cv": 0.1,
            "sampling": REF_FACTORIES["simple_pk"]}                       # defaults of model_experiment_factory: 20 samples, cv 0.1
           for k, (prior, n_t) in enumerate(itertools.product(minimal["prior_types"], minimal["timepoints"]))]
    buf = io.StringIO()
    drv.definitions_dataframe(xps).to_csv(buf, sep="\t", index=False)
    assert buf.getvalue().encode() == ref_bytes


def test_simdataset_netcdf_roundtrip(tmp_path):
    """SimDataset.to_netcdf (the reference's workflow calls dset.to_netcdf, petab_factory.py:601-602): NetCDF-3 via scipy,
    bracketed variable names, dims (sim, time)."""
    from scipy.io import netcdf_file
    ds = pkg("experiments.dataset")
    t = np.linspace(0.0, 100.0, 5)
    data = {"[y_cent]": np.arange(15.0).reshape(3, 5) / 7.0, "[y_gut]": np.exp(-np.arange(15.0).reshape(3, 5))}
    d = ds.SimDataset(t, data)
    d.to_netcdf(tmp_path / "MALE.nc")
    with netcdf_file(str(tmp_path / "MALE.nc"), "r", mmap=False) as f:
        assert f.dimensions == {"sim": 3, "time": 5}
        np.testing.assert_array_equal(f.variables["time"][:], t)
        np.testing.assert_array_equal(f.variables["sim"][:], [0, 1, 2])
        for name, values in data.items():
            assert f.variables[name].dimensions == ("sim", "time")
            np.testing.assert_array_equal(f.variables[name][:], values)
