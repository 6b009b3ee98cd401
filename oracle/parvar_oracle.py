"""CPU oracle for the parvar hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this file.  The product (``parameter-variability_b200``) never does and has no CPU fallback.

PARITY UNPINNED FOR THE ODE ROWS (the sampling / noise rows below are pinned by the reference's own code).
The arithmetic of the reference's path lives in two third-party native
libraries that are absent from /root/reference and cannot be installed here (no network):
  * libroadrunner 2.9.2 (uv.lock:961-962)  - forward simulation, SUNDIALS CVODE (BDF, Newton,
    dense LU), called at src/parvar/experiments/petab_factory.py:185-189, 229-246;
  * amici 1.0.1 (uv.lock:14-15) through pypesto 0.6.0 (uv.lock:1842-1843) - objective
    0.5*sum(log(2 pi sigma^2) + ((y-m)/sigma)^2) - log prior and its forward-sensitivity
    gradient (CVODES), called at src/parvar/optimization/petab_optimization.py:43-45,144-173.
The reference's own tests hold no golden vectors, known-answer values or tolerances for this
path (tests/experiments/test_factories.py:32 asserts a directory exists;
tests/optimization/test_optimizations.py:33-53 asserts nothing), so there is nothing to pin
against.  What this oracle is anchored on instead:
  * the ODE systems exactly as listed in src/parvar/models/sbml/*.md (restated by hand below,
    independently of the product's SBML reader / code generator);
  * exact solutions: simple_chain S1 = exp(-k1 t); simple_pk y(t) = expm(A t) y0
    (linear system), with exact sensitivities through the Van Loan block exponential;
  * for icg_body_flat, agreement of three independent integrations (scipy LSODA, Radau, BDF at
    rtol 1e-12) - checked in tests/test_oracle.py;
  * numpy's frozen legacy MT19937 stream for the sampling / noise call sites
    (petab_factory.py:101-121, 192-219), which is bit-reproducible.
PINNED: lognorm_par_transform, create_samples_parameters and add_noise are plain numpy / pandas in the reference;
tools/make_golden_reference.py executes those three definitions out of the reference file's syntax tree and commits
their outputs (tests/golden/reference_sampling_noise.npz); tests/test_golden.py holds this file's restatements (and
the product's host mirror) to them bit for bit.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np
from scipy.integrate import solve_ivp
from scipy.linalg import expm

# ----------------------------------------------------------------------------------------
# sampling / noise  (numpy legacy global RNG, petab_factory.py:94-121 and 192-219)
# ----------------------------------------------------------------------------------------


def lognorm_par_transform(loc: float, scale: float) -> tuple[float, float]:
    """(mean, sd) of a log-normal -> (mu_ln, sigma_ln).  petab_factory.py:94-98."""
    sigma_ln = np.sqrt(np.log(1.0 + (scale / loc) ** 2))
    mu_ln = np.log(loc) - 0.5 * sigma_ln**2
    return float(mu_ln), float(sigma_ln)


def create_samples_parameters(groups: dict[str, dict[str, tuple[int, float, float]]],
                              seed: Optional[int] = 1234) -> dict[str, dict[str, np.ndarray]]:
    """petab_factory.py:101-121.  ``groups[cat][par] = (n, sigma, mu)`` with the field meaning of
    ``LognormParameters`` (:71-77; filled with mu=loc, sigma=scale at :567-571).

    Quirk Q1: the reference calls ``lognorm_par_transform(s, m)`` with
    ``s = sigma`` and ``m = mu`` (:114-117), i.e. loc and scale swapped.  Reproduced literally.
    One global ``np.random.seed(seed)``, then one ``np.random.lognormal`` per (group, parameter)
    in dict order (quirk Q2).
    """
    if seed is not None:
        np.random.seed(seed)
    out: dict[str, dict[str, np.ndarray]] = {}
    for cat, pars in groups.items():
        out[cat] = {}
        for par, (n, sigma, mu) in pars.items():
            mu_ln, sigma_ln = lognorm_par_transform(sigma, mu)   # swapped on purpose (Q1)
            out[cat][par] = np.random.lognormal(mu_ln, sigma_ln, size=n)
    return out


def add_noise(values: np.ndarray, cv: float, seed: Optional[int] = None) -> np.ndarray:
    """Proportional Gaussian noise on a (T x n_obs) block.  petab_factory.py:192-219.

    ``seed = np.random.randint(0, 2001)`` drawn from, then re-seeding, the *global* stream
    (:201-204); ``errors = np.random.normal(0, 1, shape)`` row-major over (time, observable)
    (:213); ``y + y*errors*cv`` (:214-216).
    """
    if seed is None:
        seed = np.random.randint(low=0, high=2001)
    np.random.seed(seed)
    errors = np.random.normal(0, 1, values.shape)
    return values + values * errors * cv


# ----------------------------------------------------------------------------------------
# the three models, restated by hand from src/parvar/models/sbml/*.md
# ----------------------------------------------------------------------------------------
@dataclass
class OracleModel:
    name: str
    states: list[str]
    defaults: dict[str, float]                  # parameters (incl. compartment sizes)
    y0: Callable[[dict], np.ndarray]            # initial *concentrations* given parameters
    rhs: Callable[[float, Sequence, dict], list]  # works on floats and on sympy symbols
    stiff: bool = False
    obs_scale: dict[str, Callable[[dict], float]] = field(default_factory=dict)


def _chain_rhs(t, y, p):
    # simple_chain.md:24-31
    S1, S2 = y
    R1 = p["k1"] * S1
    return [-R1 / p["liver"], R1 / p["liver"]]


SIMPLE_CHAIN = OracleModel(
    name="simple_chain", states=["S1", "S2"],
    defaults={"k1": 1.0, "liver": 1.0},                       # simple_chain.md:13-17
    y0=lambda p: np.array([1.0, 0.0]),                        # simple_chain.md:19-23
    rhs=_chain_rhs)


def _pk_rhs(t, y, p):
    # simple_pk.md:29-41 (state order gut, cent, peri)
    y_gut, y_cent, y_peri = y
    ABSORPTION = p["k_abs"] * y_gut
    CLEARANCE = p["CL"] * y_cent
    R1 = p["Q"] * y_cent
    R2 = p["Q"] * y_peri
    d_cent = (ABSORPTION / p["Vcent"] - CLEARANCE / p["Vcent"] - R1 / p["Vcent"]) + R2 / p["Vcent"]
    d_gut = -ABSORPTION / p["Vgut"]
    d_peri = R1 / p["Vperi"] - R2 / p["Vperi"]
    return [d_gut, d_cent, d_peri]


SIMPLE_PK = OracleModel(
    name="simple_pk", states=["y_gut", "y_cent", "y_peri"],
    defaults={"CL": 1.0, "Q": 1.0, "Vcent": 1.0, "Vgut": 1.0, "Vperi": 1.0, "k_abs": 1.0},  # simple_pk.md:13-20
    # y_gut has initialAmount=1 (simple_pk.xml:181) -> concentration 1/Vgut; others 0 (:199,212)
    y0=lambda p: np.array([1.0 / p["Vgut"], 0.0, 0.0]),
    rhs=_pk_rhs)

ICG_STATES = ["Afeces_icg", "Car_icg", "Cgi_plasma_icg", "Chv_icg", "Cli_plasma_icg", "Clu_plasma_icg",
              "Cpo_icg", "Cre_plasma_icg", "Cve_icg", "IVDOSE_icg", "LI__icg", "LI__icg_bi"]

ICG_DEFAULTS = {  # icg_body_flat.md:17-56
    "BW": 75.0, "COBW": 1.25, "FQgi": 0.19, "FQh": 0.255, "FQlu": 1.0, "FVar": 0.0184, "FVbi": 0.00071,
    "FVgi": 0.0297, "FVhv": 0.001, "FVli": 0.021, "FVlu": 0.0076, "FVpo": 0.001, "FVve": 0.0587,
    "Fblood": 0.02, "HCT": 0.51, "HEIGHT": 170.0, "LI__ICGIM_Km": 0.021659178617926,
    "LI__ICGIM_Vmax": 0.0369598840327503, "LI__ICGIM_ki_bil": 0.02,
    "LI__ICGLI2BI_Vmax": 0.000114596604507925, "LI__ICGLI2CA_Km": 0.012388659243625,
    "LI__ICGLI2CA_Vmax": 0.000943672769975891, "LI__Vbil": 1.0, "LI__f_oatp1b3": 1.0,
    "Mr_icg": 774.96493, "Ri_icg": 0.0, "Vfeces": 1.0, "conversion_l_per_ml": 0.001,
    "conversion_ml_per_l": 1000.0, "f_bloodflow": 1.0, "f_cardiac_output": 1.0, "f_cirrhosis": 0.0,
    "f_exercise": 1.0, "f_shunts": 0.0, "f_tissue_loss": 0.0, "resection_rate": 0.0, "ti_icg": 5.0,
    # boundary species LI__bil_ext, d/dt = 0 (icg_body_flat.md:70,139)
    "LI__bil_ext": 0.01,
}


def _icg_volumes_flows(p):
    # icg_body_flat.md:79-110 (parameter-only assignment rules; BSA, ER_icg, CLinfusion_icg are outputs)
    CO = p["BW"] * p["COBW"] * p["f_cardiac_output"]
    FVre = 1 - (p["FVbi"] + p["FVgi"] + p["FVli"] + p["FVlu"] + p["FVve"] + p["FVar"] + p["FVpo"] + p["FVhv"])
    Ki_icg = (0.693 / p["ti_icg"]) * 60
    fsum = p["FVar"] + p["FVve"] + p["FVpo"] + p["FVhv"]

    def plasma_vessel(FV):
        return (1 - p["HCT"]) * (p["BW"] * FV - (FV / fsum) * p["BW"] * p["Fblood"] * (1 - fsum))

    v = {}
    v["Var"] = plasma_vessel(p["FVar"])
    v["Vbi"] = p["BW"] * p["FVbi"]
    Vgi = p["BW"] * p["FVgi"]
    v["Vhv"] = plasma_vessel(p["FVhv"])
    Vli = p["BW"] * p["FVli"] * (1 - p["resection_rate"])
    Vlu = p["BW"] * p["FVlu"]
    v["Vpo"] = plasma_vessel(p["FVpo"])
    v["Vve"] = plasma_vessel(p["FVve"])
    QC = (CO / 1000) * 60
    v["Vgi_plasma"] = Vgi * p["Fblood"] * (1 - p["HCT"])
    v["Vli_plasma"] = Vli * p["Fblood"] * (1 - p["HCT"])
    v["Vli_tissue"] = Vli * (1 - p["f_tissue_loss"]) * (1 - p["Fblood"])
    v["Vlu_plasma"] = Vlu * p["Fblood"] * (1 - p["HCT"])
    Vre = p["BW"] * FVre
    v["Vre_plasma"] = Vre * p["Fblood"] * (1 - p["HCT"])
    v["Ki_icg"] = Ki_icg
    v["Qgi"] = QC * p["FQgi"] * p["f_bloodflow"]
    v["Qh"] = QC * p["FQh"] * p["f_bloodflow"] * p["f_exercise"]
    v["Qlu"] = QC * p["FQlu"]
    FQre = 1 - v["Qh"] / v["Qlu"]
    v["Qpo"] = v["Qgi"]
    v["Qha"] = v["Qh"] - v["Qpo"]
    v["Qre"] = QC * FQre
    return v


def _icg_rhs(t, y, p):
    # icg_body_flat.md:77-141, states in the listing's (alphabetical) order
    (Afeces, Car, Cgi, Chv, Cli, Clu, Cpo, Cre, Cve, IVDOSE, LIicg, LIicg_bi) = y
    v = _icg_volumes_flows(p)
    iv_icg = v["Ki_icg"] * IVDOSE / p["Mr_icg"]
    ICGIM = (p["LI__f_oatp1b3"] * p["LI__ICGIM_Vmax"] * v["Vli_tissue"] * Cli
             / (p["LI__ICGIM_Km"] * (1 + p["LI__bil_ext"] / p["LI__ICGIM_ki_bil"]) + Cli))
    ICGLI2BI = p["LI__ICGLI2BI_Vmax"] * v["Vli_tissue"] * LIicg_bi
    ICGLI2CA = p["LI__ICGLI2CA_Vmax"] * v["Vli_tissue"] * LIicg / (p["LI__ICGLI2CA_Km"] + LIicg)
    Flow_ar_argi = v["Qgi"] * Car
    Flow_gi_po = v["Qgi"] * Cgi
    Flow_hv_ve = v["Qh"] * Chv
    Flow_lu_ar = v["Qlu"] * Clu
    Flow_ve_lu = v["Qlu"] * Cve
    Flow_po_hv = p["f_shunts"] * v["Qpo"] * Cpo
    Flow_po_li = (1 - p["f_shunts"]) * v["Qpo"] * Cpo
    Flow_ar_arre = v["Qre"] * Car
    Flow_arli_hv = p["f_shunts"] * v["Qha"] * Car
    Flow_arli_li = (1 - p["f_shunts"]) * v["Qha"] * Car
    Flow_li_hv = (1 - p["f_shunts"]) * (v["Qpo"] + v["Qha"]) * Cli
    Flow_re_ve = v["Qre"] * Cre
    Var, Vhv, Vpo, Vve = v["Var"], v["Vhv"], v["Vpo"], v["Vve"]
    return [
        ICGLI2BI,
        (-Flow_ar_arre / Var - Flow_ar_argi / Var - Flow_arli_li / Var - Flow_arli_hv / Var) + Flow_lu_ar / Var,
        Flow_ar_argi / v["Vgi_plasma"] - Flow_gi_po / v["Vgi_plasma"],
        Flow_arli_hv / Vhv + Flow_po_hv / Vhv + Flow_li_hv / Vhv - Flow_hv_ve / Vhv,
        (Flow_arli_li / v["Vli_plasma"] + Flow_po_li / v["Vli_plasma"] - Flow_li_hv / v["Vli_plasma"]
         - ICGIM / v["Vli_plasma"]),
        Flow_ve_lu / v["Vlu_plasma"] - Flow_lu_ar / v["Vlu_plasma"],
        Flow_gi_po / Vpo - Flow_po_li / Vpo - Flow_po_hv / Vpo,
        Flow_ar_arre / v["Vre_plasma"] - Flow_re_ve / v["Vre_plasma"],
        iv_icg / Vve + Flow_re_ve / Vve + Flow_hv_ve / Vve - Flow_ve_lu / Vve,
        -iv_icg * p["Mr_icg"] + p["Ri_icg"],
        ICGIM / v["Vli_tissue"] - ICGLI2CA / v["Vli_tissue"],
        ICGLI2CA / v["Vbi"] - ICGLI2BI / v["Vbi"],
    ]


ICG_BODY_FLAT = OracleModel(
    name="icg_body_flat", states=ICG_STATES, defaults=ICG_DEFAULTS,
    y0=lambda p: np.zeros(12),                                # icg_body_flat.md:60-74
    rhs=_icg_rhs, stiff=True,
    # "[Afeces_icg]" = amount / Vfeces (species with hasOnlySubstanceUnits, icg_body_flat.xml)
    obs_scale={"Afeces_icg": lambda p: 1.0 / p["Vfeces"]})

MODELS: dict[str, OracleModel] = {m.name: m for m in (SIMPLE_CHAIN, SIMPLE_PK, ICG_BODY_FLAT)}


# ----------------------------------------------------------------------------------------
# forward simulation:  RoadRunner.resetAll / setValue / simulate(start, end, steps)
# (petab_factory.py:229-246) on the uniform grid linspace(start, end, steps+1)
# ----------------------------------------------------------------------------------------
def _params(model: OracleModel, overrides: Optional[dict]) -> tuple[dict, dict]:
    p = dict(model.defaults)
    y0_over = {}
    for k, v in (overrides or {}).items():
        key = k[1:-1] if k.startswith("[") and k.endswith("]") else k
        if key in model.states:
            y0_over[key] = float(v)
        elif key in p:
            p[key] = float(v)
        else:
            raise KeyError(f"{model.name}: unknown id {k!r}")
    return p, y0_over


def simulate(model: OracleModel | str, overrides: Optional[dict] = None,
             tgrid: Sequence[float] = (), rtol: float = 1e-12, atol: float = 1e-14,
             method: Optional[str] = None) -> np.ndarray:
    """One trajectory on ``tgrid`` -> array [T, n_states] of concentrations ("[sid]" columns).

    ``overrides`` plays the role of the ``setValue`` calls (dosage dict first, then the
    parameter row; :231-239): parameter ids change parameters, state ids change initial values.
    """
    if isinstance(model, str):
        model = MODELS[model]
    p, y0_over = _params(model, overrides)
    y0 = np.array(model.y0(p), dtype=float)
    for k, v in y0_over.items():
        y0[model.states.index(k)] = v
    tgrid = np.asarray(tgrid, dtype=float)
    method = method or ("LSODA" if model.stiff else "DOP853")
    t0 = float(tgrid[0])
    if len(tgrid) == 1 or tgrid[-1] == t0:
        return np.repeat(y0[None, :], len(tgrid), axis=0)
    sol = solve_ivp(lambda t, y: model.rhs(t, y, p), (t0, float(tgrid[-1])), y0, method=method,
                    t_eval=tgrid, rtol=rtol, atol=atol)
    if not sol.success:
        raise RuntimeError(f"oracle integration failed: {sol.message}")
    Y = sol.y.T.copy()
    for sid, fn in model.obs_scale.items():
        Y[:, model.states.index(sid)] *= fn(p)
    return Y


def simulate_reference_loop(model: OracleModel | str, rows: Sequence[dict], start: float, end: float,
                            steps: int, dosage: Optional[dict] = None, observables: Optional[Sequence[str]] = None,
                            rtol: float = 1e-6, atol: float = 1e-6) -> np.ndarray:
    """Literal restatement of the hot loop of ``simulate_samples`` (petab_factory.py:221-272, without
    noise): one solve per parameter row, the reference's tolerances abs=rel=1e-6 (:183-189, quirk
    Q9), LSODA standing in for CVODE's BDF/Adams.  Returns [n, T, n_obs]."""
    if isinstance(model, str):
        model = MODELS[model]
    tgrid = np.linspace(start, end, steps + 1)
    out = []
    for row in rows:
        over = dict(dosage or {})
        over.update(row)
        Y = simulate(model, over, tgrid, rtol=rtol, atol=atol, method="LSODA")
        if observables:
            Y = Y[:, [model.states.index(o.strip("[]")) for o in observables]]
        out.append(Y)
    return np.stack(out)


# ----------------------------------------------------------------------------------------
# exact solutions for the linear models
# ----------------------------------------------------------------------------------------
def simple_chain_exact(k1: float, tgrid: Sequence[float], liver: float = 1.0) -> np.ndarray:
    t = np.asarray(tgrid, dtype=float)
    k = k1 / liver
    return np.stack([np.exp(-k * t), -np.expm1(-k * t)], axis=1)


def simple_chain_exact_sens(k1: float, tgrid: Sequence[float]) -> np.ndarray:
    """d[S1,S2]/dk1 -> [T, 2, 1]."""
    t = np.asarray(tgrid, dtype=float)
    return np.stack([-t * np.exp(-k1 * t), t * np.exp(-k1 * t)], axis=1)[:, :, None]


def simple_pk_matrix(p: dict) -> np.ndarray:
    """A with dy/dt = A y, order (gut, cent, peri); from simple_pk.md:29-41."""
    k, CL, Q = p["k_abs"], p["CL"], p["Q"]
    Vg, Vc, Vp = p["Vgut"], p["Vcent"], p["Vperi"]
    return np.array([[-k / Vg, 0.0, 0.0],
                     [k / Vc, -(CL + Q) / Vc, Q / Vc],
                     [0.0, Q / Vp, -Q / Vp]])


def simple_pk_exact(overrides: Optional[dict], tgrid: Sequence[float]) -> np.ndarray:
    p, y0_over = _params(SIMPLE_PK, overrides)
    y0 = SIMPLE_PK.y0(p)
    for k, v in y0_over.items():
        y0[SIMPLE_PK.states.index(k)] = v
    A = simple_pk_matrix(p)
    return np.stack([expm(A * float(t)) @ y0 for t in tgrid])


def simple_pk_exact_sens(overrides: Optional[dict], tgrid: Sequence[float],
                         sens: Sequence[str] = ("k_abs", "CL")) -> np.ndarray:
    """d y / d theta via expm([[A, dA],[0, A]] t): the top-right block times y0.  -> [T, 3, n_sens]."""
    p, y0_over = _params(SIMPLE_PK, overrides)
    y0 = SIMPLE_PK.y0(p)
    for k, v in y0_over.items():
        y0[SIMPLE_PK.states.index(k)] = v
    A = simple_pk_matrix(p)
    out = np.zeros((len(tgrid), 3, len(sens)))
    for si, name in enumerate(sens):
        eps = 1e-6 * max(abs(p[name]), 1.0)
        pp, pm = dict(p), dict(p)
        pp[name] += eps
        pm[name] -= eps
        dA = (simple_pk_matrix(pp) - simple_pk_matrix(pm)) / (2 * eps)   # A is linear in k_abs, CL, Q
        M = np.zeros((6, 6))
        M[:3, :3] = A
        M[:3, 3:] = dA
        M[3:, 3:] = A
        for ti, t in enumerate(tgrid):
            out[ti, :, si] = expm(M * float(t))[:3, 3:] @ y0
    return out


# ----------------------------------------------------------------------------------------
# forward sensitivities for any model: sympy Jacobians of the hand-written RHS, integrated
# together with the states (what CVODES does for AMICI; petab_optimization.py:144-151)
# ----------------------------------------------------------------------------------------
_SENS_CACHE: dict = {}


def _sens_functions(model: OracleModel, sens: tuple[str, ...]):
    key = (model.name, sens)
    if key in _SENS_CACHE:
        return _SENS_CACHE[key]
    import sympy as sp
    ys = sp.symbols([f"y_{s}" for s in model.states])
    ps = {k: sp.Symbol(f"p_{k}") for k in model.defaults}
    f = sp.Matrix(model.rhs(sp.Symbol("t"), ys, ps))
    J = f.jacobian(ys)
    Fp = f.jacobian([ps[s] for s in sens])
    args = [ys, [ps[k] for k in model.defaults]]
    fJ = sp.lambdify(args, J, "numpy", cse=True)
    fP = sp.lambdify(args, Fp, "numpy", cse=True)
    _SENS_CACHE[key] = (fJ, fP)
    return fJ, fP


def simulate_sens(model: OracleModel | str, overrides: Optional[dict], tgrid: Sequence[float],
                  sens: Sequence[str], rtol: float = 1e-12, atol: float = 1e-14,
                  method: Optional[str] = None) -> tuple[np.ndarray, np.ndarray]:
    """-> (Y [T, n], S [T, n, n_sens]) with S = dY/dtheta (concentration outputs)."""
    if isinstance(model, str):
        model = MODELS[model]
    sens = tuple(sens)
    p, y0_over = _params(model, overrides)
    n, ns = len(model.states), len(sens)
    fJ, fP = _sens_functions(model, sens)
    pvec = [p[k] for k in model.defaults]
    y0 = np.array(model.y0(p), dtype=float)
    # d y0 / d theta by central differences of the (closed-form) initial state
    s0 = np.zeros((n, ns))
    for si, name in enumerate(sens):
        eps = 1e-6 * max(abs(p[name]), 1e-3)
        pp, pm = dict(p), dict(p)
        pp[name] += eps
        pm[name] -= eps
        s0[:, si] = (np.array(model.y0(pp)) - np.array(model.y0(pm))) / (2 * eps)
    for k, v in y0_over.items():
        y0[model.states.index(k)] = v
        s0[model.states.index(k), :] = 0.0

    def aug(t, z):
        y = z[:n]
        S = z[n:].reshape(n, ns)
        dy = np.array(model.rhs(t, y, p), dtype=float)
        J = np.asarray(fJ(y, pvec), dtype=float)
        P = np.asarray(fP(y, pvec), dtype=float)
        return np.concatenate([dy, (J @ S + P).ravel()])

    tgrid = np.asarray(tgrid, dtype=float)
    method = method or ("LSODA" if model.stiff else "DOP853")
    sol = solve_ivp(aug, (float(tgrid[0]), float(tgrid[-1])), np.concatenate([y0, s0.ravel()]),
                    method=method, t_eval=tgrid, rtol=rtol, atol=atol)
    if not sol.success:
        raise RuntimeError(sol.message)
    Z = sol.y.T
    Y = Z[:, :n].copy()
    S = Z[:, n:].reshape(len(tgrid), n, ns).copy()
    for sid, fn in model.obs_scale.items():
        i = model.states.index(sid)
        Y[:, i] *= fn(p)
        S[:, i, :] *= fn(p)
    return Y, S


# ----------------------------------------------------------------------------------------
# likelihood and gradient
# ----------------------------------------------------------------------------------------
LOG_2PI = math.log(2.0 * math.pi)


def loglik_normal(sim: np.ndarray, data: np.ndarray, sigma) -> float:
    """log N(data; sim, sigma^2) summed.  AMICI's objective is the negative of this:
    nllh = 0.5*sum(log(2 pi sigma^2) + ((y - m)/sigma)^2)  [amici 1.0.1, normal noise, lin scale];
    PEtab tables written at petab_factory.py:388-398 fix sigma = 1 (quirk Q5).
    ``data`` may have leading individual axes (pooled fit, quirk Q7): it broadcasts against sim."""
    sigma = np.broadcast_to(np.asarray(sigma, dtype=float), np.broadcast_shapes(np.shape(sim), np.shape(data)))
    r = (data - sim) / sigma
    return float(-0.5 * np.sum(np.log(2 * np.pi * sigma**2) + r**2))


def loglik_normal_grad(sim: np.ndarray, dsim: np.ndarray, data: np.ndarray, sigma) -> np.ndarray:
    """d loglik / d theta = sum (y - f)/sigma^2 * df/dtheta.  ``dsim`` = sim.shape + (n_sens,)."""
    sigma = np.asarray(sigma, dtype=float)
    w = (data - sim) / sigma**2
    w = np.broadcast_to(w, np.broadcast_shapes(w.shape, sim.shape))
    if w.ndim > sim.ndim:                      # pooled: sum the individual axes first
        w = w.reshape((-1,) + sim.shape).sum(axis=0)
    return np.tensordot(w, dsim, axes=sim.ndim)


def loglik_lognormal(sim: np.ndarray, data: np.ndarray, sigma) -> float:
    """log-normal observation model (BASELINE.json north_star; PEtab 'log-normal' noise on a lin
    observable): log p = -log(y sigma sqrt(2 pi)) - (log y - log f)^2 / (2 sigma^2)."""
    sigma = np.broadcast_to(np.asarray(sigma, dtype=float), np.broadcast_shapes(np.shape(sim), np.shape(data)))
    r = (np.log(data) - np.log(sim)) / sigma
    return float(np.sum(-np.log(data) - np.log(sigma) - 0.5 * LOG_2PI - 0.5 * r**2))


def loglik_lognormal_grad(sim: np.ndarray, dsim: np.ndarray, data: np.ndarray, sigma) -> np.ndarray:
    sigma = np.asarray(sigma, dtype=float)
    w = (np.log(data) - np.log(sim)) / sigma**2 / sim
    w = np.broadcast_to(w, np.broadcast_shapes(w.shape, sim.shape))
    if w.ndim > sim.ndim:
        w = w.reshape((-1,) + sim.shape).sum(axis=0)
    return np.tensordot(w, dsim, axes=sim.ndim)


def log_prior_normal(x: np.ndarray, loc: np.ndarray, scale: np.ndarray) -> tuple[float, np.ndarray]:
    """PEtab ``parameterScaleNormal`` prior on lin-scale parameters (petab_factory.py:443-449, quirk Q6):
    sum log N(x; loc, scale); gradient -(x-loc)/scale^2."""
    x, loc, scale = (np.asarray(a, dtype=float) for a in (x, loc, scale))
    lp = -0.5 * np.sum(np.log(2 * np.pi * scale**2) + ((x - loc) / scale) ** 2)
    return float(lp), -(x - loc) / scale**2


def population_lognormal_logpdf(theta: np.ndarray, mu_ln: np.ndarray, sigma_ln: np.ndarray):
    """log LogNormal(theta_i; mu_ln, sigma_ln) summed over individuals, with gradients wrt theta_i,
    mu_ln, sigma_ln (the hierarchical form of BASELINE.json configs 3-4; build-defined because the
    reference has no hierarchy).  theta [n_ind, n_par]; mu_ln, sigma_ln [n_par]."""
    theta = np.asarray(theta, dtype=float)
    z = (np.log(theta) - mu_ln) / sigma_ln
    lp = np.sum(-np.log(theta) - np.log(sigma_ln) - 0.5 * LOG_2PI - 0.5 * z**2)
    d_theta = (-1.0 - z / sigma_ln) / theta
    d_mu = np.sum(z / sigma_ln, axis=0)
    d_sigma = np.sum((z**2 - 1.0) / sigma_ln, axis=0)
    return float(lp), d_theta, d_mu, d_sigma
