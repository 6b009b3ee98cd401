#!/usr/bin/env python
"""Generate tests/golden/*.npz - golden input/output vectors for the hot path.

The reference's own implementation (libroadrunner / AMICI behind parvar) cannot be imported in this
container (Python 3.12 < required 3.13, none of the native dependencies installed,
no network), so these vectors come from the ORACLE (oracle/parvar_oracle.py: exact solutions for the
linear models, scipy LSODA at rtol 1e-12 for icg_body_flat, sympy forward sensitivities), seeded like
the reference's sampling call site (np.random.seed(1234), petab_factory.py:101-118).  They pin the oracle
against regression and give the CUDA path fixed vectors that do not depend on scipy at test time.

usage: python tools/make_golden.py     (rewrites tests/golden/)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import parvar_oracle as O  # noqa: E402

OUT = ROOT / "tests" / "golden"
MU, SG = -0.3465735902799726, 0.8325546111576977      # lognorm_par_transform(1, 1)


def simple_pk():
    B = 64
    np.random.seed(1234)
    k_abs = np.random.lognormal(MU, SG, B)
    CL = np.random.lognormal(MU, SG, B)
    t = np.linspace(0, 100, 50)
    Y = np.stack([O.simple_pk_exact({"k_abs": a, "CL": c}, t) for a, c in zip(k_abs, CL)], axis=-1)   # [T,3,B]
    S = np.stack([O.simple_pk_exact_sens({"k_abs": a, "CL": c}, t) for a, c in zip(k_abs, CL)], axis=-1)  # [T,3,2,B]
    truth = O.simple_pk_exact({}, t)
    rs = np.random.RandomState(1234)
    data = truth * (1.0 + 0.05 * rs.normal(size=truth.shape))          # C1's data vector
    sigma = 1.0
    logp = np.array([O.loglik_normal(Y[:, :, b], data, sigma) for b in range(B)])
    grad = np.stack([O.loglik_normal_grad(Y[:, :, b], S[:, :, :, b], data, sigma) for b in range(B)], axis=-1)
    np.savez_compressed(OUT / "simple_pk_c2.npz", k_abs=k_abs, CL=CL, t=t, Y=Y, S=S, data=data, sigma=sigma,
                        logp=logp, grad=grad, states=np.array(O.SIMPLE_PK.states))


def simple_chain():
    B = 32
    np.random.seed(1234)
    k1 = np.random.lognormal(MU, SG, B)
    t = np.linspace(0, 100, 21)
    Y = np.stack([O.simple_chain_exact(k, t) for k in k1], axis=-1)
    S = np.stack([O.simple_chain_exact_sens(k, t) for k in k1], axis=-1)
    np.savez_compressed(OUT / "simple_chain.npz", k1=k1, t=t, Y=Y, S=S, states=np.array(O.SIMPLE_CHAIN.states))


def icg():
    B = 16
    np.random.seed(1234)
    # C4 inputs: (mean, sd) = (75, 10) and (0.03696, 0.01), unswapped
    m1, s1 = O.lognorm_par_transform(75.0, 10.0)
    m2, s2 = O.lognorm_par_transform(0.0369598840327503, 0.01)
    BW = np.random.lognormal(m1, s1, B)
    Vmax = np.random.lognormal(m2, s2, B)
    t = np.linspace(0, 100, 21)
    Ys, Ss = [], []
    for b in range(B):
        over = {"IVDOSE_icg": 10.0, "BW": BW[b], "LI__ICGIM_Vmax": Vmax[b]}
        if b < 4:
            Yb, Sb = O.simulate_sens("icg_body_flat", over, t, ["BW", "LI__ICGIM_Vmax"])
            Ss.append(Sb)
        else:
            Yb = O.simulate("icg_body_flat", over, t)
        Ys.append(Yb)
    Y = np.stack(Ys, axis=-1)                 # [T, 12, B]
    S = np.stack(Ss, axis=-1)                 # [T, 12, 2, 4]
    obs = ["Cre_plasma_icg", "Cgi_plasma_icg"]
    idx = [O.ICG_STATES.index(o) for o in obs]
    truth = O.simulate("icg_body_flat", {"IVDOSE_icg": 10.0}, t)[:, idx]
    rs = np.random.RandomState(1234)
    data = truth * (1.0 + 0.05 * rs.normal(size=truth.shape))
    sigma = 1e-4
    logp = np.array([O.loglik_normal(Y[..., b][:, idx], data, sigma) for b in range(B)])
    grad = np.stack([O.loglik_normal_grad(Y[..., b][:, idx], S[..., b][:, idx, :], data, sigma) for b in range(4)], axis=-1)
    np.savez_compressed(OUT / "icg_body_flat_c4.npz", BW=BW, Vmax=Vmax, t=t, Y=Y, S=S, data=data, sigma=sigma,
                        logp=logp, grad=grad, states=np.array(O.ICG_STATES), obs=np.array(obs))


def sampling():
    """The reference's sampling / noise stream (petab_factory.py:101-121, 192-219) for factories/simple_pk.py:19-44."""
    s = O.create_samples_parameters({
        "MALE": {"CL": (20, 1.0, 1.0), "k_abs": (20, 1.0, 0.5)},
        "FEMALE": {"CL": (20, 1.0, 0.5), "k_abs": (20, 1.0, 1.0)}})
    y = np.linspace(0, 1, 63).reshape(21, 3)
    noisy = O.add_noise(y, 0.05)          # continues the global stream right after the sampling calls
    np.savez_compressed(OUT / "sampling_simple_pk.npz", CL_MALE=s["MALE"]["CL"], k_abs_MALE=s["MALE"]["k_abs"],
                        CL_FEMALE=s["FEMALE"]["CL"], k_abs_FEMALE=s["FEMALE"]["k_abs"], y=y, noisy=noisy)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    simple_pk()
    simple_chain()
    icg()
    sampling()
    for f in sorted(OUT.glob("*.npz")):
        print(f, f.stat().st_size)
