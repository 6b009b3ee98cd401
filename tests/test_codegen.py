"""SBML reader + code generator against the oracle's hand-written restatement of the models."""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import pkg
from host import harness

REF_SBML = Path("/root/reference/src/parvar/models/sbml")


def _theta(meta, over=None):
    th = np.array(meta["theta_default"], float)
    for k, v in (over or {}).items():
        th[meta["theta"].index(k)] = v
    return th


@pytest.mark.parametrize("model", ["simple_chain", "simple_pk", "icg_body_flat"])
def test_generated_rhs_matches_oracle(model, oracle):
    om = oracle.MODELS[model]
    meta = harness.meta(model)
    rs = np.random.RandomState(3)
    for trial in range(10):
        over = {}
        for k in meta["sens"]:
            over[k] = om.defaults[k] * rs.lognormal(0, 0.3)
        p = dict(om.defaults)
        p.update(over)
        y_or = rs.rand(len(om.states)) * 1e-2
        if model == "icg_body_flat":
            y_or[om.states.index("IVDOSE_icg")] = rs.rand() * 10
        y_gen = np.array([y_or[om.states.index(s)] for s in meta["states"]])
        f_gen, x, Jv = harness.rhs_jac(model, _theta(meta, over), y_gen)
        f_or = np.array(om.rhs(0.0, y_or, p), dtype=float)
        f_or = np.array([f_or[om.states.index(s)] for s in meta["states"]])
        np.testing.assert_allclose(f_gen, f_or, rtol=1e-11, atol=1e-14 * np.abs(f_or).max())
        # Jacobian on its sparse pattern vs central differences of the oracle RHS
        n = len(om.states)
        J = np.zeros((n, n))
        for (i, j), v in zip(meta["jac_pattern"], Jv):
            J[i, j] = v
        Jfd = np.zeros((n, n))
        for j, sj in enumerate(meta["states"]):
            e = 1e-7 * max(abs(y_gen[j]), 1e-3)
            yp, ym = y_or.copy(), y_or.copy()
            yp[om.states.index(sj)] += e
            ym[om.states.index(sj)] -= e
            d = (np.array(om.rhs(0, yp, p), float) - np.array(om.rhs(0, ym, p), float)) / (2 * e)
            Jfd[:, j] = [d[om.states.index(s)] for s in meta["states"]]
        np.testing.assert_allclose(J, Jfd, rtol=2e-5, atol=1e-7 * np.abs(Jfd).max())
        # generated sparse LU: (3 I - J) x = f
        np.testing.assert_allclose((3 * np.eye(n) - J) @ x, f_gen, rtol=1e-9, atol=1e-12 * np.abs(f_gen).max())


def test_model_structure():
    cg = pkg("codegen")
    P = pkg()
    sb = cg.read_sbml(P.MODELS["icg_body_flat"])
    ode = cg.build_ode_system(sb)
    assert len(ode.states) == 12 and "IVDOSE_icg" in ode.states and "LI__bil_ext" in ode.theta
    assert set(ode.dead_rules) >= {"BSA", "ER_icg", "CLinfusion_icg"}
    hs = cg.hoist_system(ode, ["BW", "LI__ICGIM_Vmax"])
    assert sum(1 for v in hs.jac_const.values() if not v) == 4             # only the Michaelis-Menten entries move
    sb = cg.read_sbml(P.MODELS["simple_pk"])
    ode = cg.build_ode_system(sb)
    assert ode.states == ["y_gut", "y_cent", "y_peri"] and ode.theta[:3] == ["k_abs", "CL", "Q"]
    with pytest.raises(ValueError):
        cg.hoist_system(ode, ["not_a_parameter"])


@pytest.mark.skipif(not REF_SBML.exists(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("model", ["simple_chain", "simple_pk", "icg_body_flat"])
def test_math_only_copy_equals_reference_sbml(model):
    """models/sbml/*.xml (tools/strip_sbml.py) describe exactly the reference's ODE systems."""
    cg = pkg("codegen")
    P = pkg()
    a = cg.build_ode_system(cg.read_sbml(P.MODELS[model]))
    b = cg.build_ode_system(cg.read_sbml(REF_SBML / f"{model}.xml"))
    a.name = b.name = model
    sens = cg.BUILTIN[model]
    assert cg.model_fingerprint(a, sens) == cg.model_fingerprint(b, sens)
    meta = json.loads((Path(P.PKG_DIR) / "csrc" / "generated" / f"{model}.json").read_text())
    assert meta["fingerprint"] == cg.model_fingerprint(b, sens)


def test_generated_sources_are_current(tmp_path):
    """csrc/generated/* is exactly what the generator emits from models/sbml/* (no hand edits, no drift)."""
    cg = pkg("codegen")
    P = pkg()
    out = cg.generate_builtin(outdir=tmp_path)
    gen = Path(P.PKG_DIR) / "csrc" / "generated"
    for name in list(out) + ["model_tables"]:
        ext = ".inc" if name == "model_tables" else ".cuh"
        assert (tmp_path / f"{name}{ext}").read_text() == (gen / f"{name}{ext}").read_text()


def test_unsupported_sbml_is_rejected(tmp_path):
    cg = pkg("codegen")
    bad = tmp_path / "bad.xml"
    bad.write_text('<sbml xmlns="http://www.sbml.org/sbml/level3/version2/core" level="3" version="2"><model id="m">'
                   '<listOfEvents><event id="e"/></listOfEvents></model></sbml>')
    with pytest.raises(NotImplementedError):
        cg.read_sbml(bad)


def test_no_ieee_division_inside_a_step():
    """The per-stage functions and the LU factorisation of the generated headers multiply by pv_rcp_newton(d) instead of
    dividing: an IEEE fp64 division carries a slow-path branch that splits the straight-line step.
    hoist / hoist_sens run once per trajectory and keep the IEEE division."""
    import re
    gen = pkg("build").CSRC / "generated"
    for name in ("simple_chain", "simple_pk", "icg_body_flat"):
        text = (gen / f"{name}.cuh").read_text()
        for fn in ("rhs", "rhs_sens", "jac", "lu_factor", "lu_solve", "sens_stage_term", "sens_rhs_block", "sens_term_prep",
                   "sens_term_apply"):
            m = re.search(r"PV_HD static void " + fn + r"\(.*?\n    \}\n", text, flags=re.S)
            assert m, (name, fn)
            body = re.sub(r"//.*", "", m.group(0))
            body = re.sub(r"\(\d+\.0/\d+\.0\)", "", body)       # rational constants fold at compile time
            assert "/" not in body, (name, fn)
    icg = (gen / "icg_body_flat.cuh").read_text()
    assert icg.count("pv_rcp_newton") >= 12 + 4          # 12 pivots + the Michaelis-Menten denominators


def perturbed_simple_pk(tmp_path):
    """simple_pk with an extra, saturable elimination from the peripheral compartment (two new parameters): an ODE system
    that is in no registry - and not linear in its coefficients, unlike the three shipped models' explicit ones."""
    P = pkg()
    xml = P.MODELS["simple_pk"].read_text()
    xml = xml.replace('<parameter id="Q" value="1" constant="true" />',
                      '<parameter id="Q" value="1" constant="true" />\n   <parameter id="k_el" value="0.3" constant="true" />\n'
                      '   <parameter id="Km" value="0.05" constant="true" />')
    reaction = '''   <reaction id="ELIMINATION" reversible="false">
    <listOfReactants>
     <speciesReference species="y_peri" stoichiometry="1" constant="true" />
    </listOfReactants>
    <kineticLaw>
     <math:math>
      <math:apply>
       <math:divide />
       <math:apply><math:times /><math:ci>k_el</math:ci><math:ci>y_peri</math:ci></math:apply>
       <math:apply><math:plus /><math:ci>Km</math:ci><math:ci>y_peri</math:ci></math:apply>
      </math:apply>
     </math:math>
    </kineticLaw>
   </reaction>
  </listOfReactions>'''
    xml = xml.replace("  </listOfReactions>", reaction).replace('<model id="simple_pk">', '<model id="pk_saturable">')
    path = tmp_path / "pk_saturable.xml"
    path.write_text(xml)
    return path


def test_runtime_compile_of_an_unregistered_model(tmp_path):
    """codegen.compile_model: an SBML file outside the registry becomes its own build of the library (same C ABI, one
    model, the requested sensitivity set), cached by fingerprint - the reference's roadrunner JIT / AMICI model build
    (petab_factory.py:183-190, run_optimization.py:43-46).  No GPU needed to build, load and inspect it."""
    cg = pkg("codegen")
    rt = pkg("runtime")
    pf = pkg("experiments.petab_factory")
    path = perturbed_simple_pk(tmp_path)
    with pytest.raises(rt.ParvarB200Error, match="not compiled"):
        pf.resolve_model(path, compile_unknown=False)
    name, lib_path = cg.compile_model(path, sens=["k_abs", "k_el"], cache_dir=tmp_path / "cache")
    assert name == "pk_saturable" and lib_path.exists()
    again = cg.compile_model(path, sens=["k_abs", "k_el"], cache_dir=tmp_path / "cache")
    assert again == (name, lib_path)                                   # cached: same fingerprint, no second build
    other = cg.compile_model(path, sens=["CL"], cache_dir=tmp_path / "cache")[1]
    assert other != lib_path                                           # another sensitivity set = another build
    lib = rt.load_model_library("pk_saturable@test", lib_path)
    assert lib.pv_model_count() == 1 and lib.pv_model_name(0) == b"pk_saturable"
    meta = json.loads(lib_path.with_suffix(".json").read_text())
    assert meta["states"] == ["y_gut", "y_cent", "y_peri"] and {"k_el", "Km"} <= set(meta["theta"])
    assert meta["sens"] == ["k_abs", "k_el"]
    header = cg.generate_model(path, ["k_abs", "k_el"]).header
    assert "RHS_LINEAR_IN_C = false" in header                         # Km sits inside a denominator
    if lib.pv_device_count() <= 0:                                     # still no CPU fallback in a run-time build
        assert not lib.pv_problem_create(b"pk_saturable", 0) and b"no CUDA device" in lib.pv_last_error()
