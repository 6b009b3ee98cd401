"""Host-side mirror of the objective the reference's optimiser and sampler call.

In the reference, ``PyPestoSampler.load_problem`` builds ``pypesto_problem.objective`` from the PEtab
tables through AMICI (src/parvar/optimization/petab_optimization.py:41-48) and
``pypesto.optimize.minimize`` / ``pypesto.sample.sample`` call it as ``objective(x, sensi_orders)``
(:144-151, :168-173).  ``PetabObjective`` is that callable with the same conventions:

  * value = NEGATIVE log posterior (minimised):  0.5*sum(log(2 pi sigma^2) + ((y-f)/sigma)^2) - log prior;
  * ``objective(x)`` -> fval; ``objective(x, sensi_orders=(0, 1))`` -> (fval, grad);
    ``objective(x, sensi_orders=(1,))`` -> grad; ``return_dict=True`` -> {"fval":, "grad":};
  * x ordered like the PEtab parameter table the reference writes: ``for par in params: for group in
    groups: f"{par}_{group}"`` (src/parvar/experiments/petab_factory.py:428-451), bounds [0, 1e8] (:437-438);
  * a failed integration gives fval = inf / grad = nan (pyPESTO treats non-finite values as failed
    evaluations).

plus the batched entry ``logp_and_grad(X[B, dim])`` for chains / multistarts, which
is what makes the path data parallel: one launch integrates B x n_groups trajectories.

``pypesto.Problem(pypesto.Objective(fun=obj.fun, grad=obj.grad), lb=obj.lb, ub=obj.ub,
x_names=obj.x_names)`` plugs it into the reference's unchanged optimiser / sampler code.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from .. import runtime

LOG_2PI = float(np.log(2.0 * np.pi))


@dataclass
class GroupData:
    """Measurements of one group (PEtab simulation condition).  ``y`` [n_individuals, T, n_obs];
    NaN = missing.  All individuals share the group's parameters (pooled fit, petab_factory.py:409)."""
    id: str
    y: np.ndarray


def sufficient_statistics(y: np.ndarray, noise_model: int) -> np.ndarray:
    """[n_ind, T, K] measurements -> [3, T, K] (n, sum, sum of squares) of y or log y."""
    y = np.asarray(y, dtype=np.float64)
    v = np.log(y) if noise_model == runtime.NOISE_LOGNORMAL else y
    ok = np.isfinite(v)
    v0 = np.where(ok, v, 0.0)
    return np.stack([ok.sum(axis=0).astype(np.float64), v0.sum(axis=0), (v0 * v0).sum(axis=0)])


class PetabObjective:
    """Negative log posterior of a parvar PEtab problem, evaluated on the GPU.

    Parameters
    ----------
    model:        registry name ("simple_chain", "simple_pk", "icg_body_flat")
    parameters:   estimated model parameters, e.g. ["k_abs", "CL"]; must be the model's compiled
                  sensitivity parameters (in any order) when gradients are requested
    groups:       one ``GroupData`` per group, in PEtab condition order
    timepoints:   measurement times [T]
    observables:  observed species ids [K] (with or without brackets)
    sigma:        noise sd (scalar or [K]); the reference writes ``noiseFormula = 1`` (:394)
    priors:       {x_name: (loc, scale)} -> ``parameterScaleNormal`` prior on the lin scale (:443-449)
    mode:         "intended": SBML initial values + ``dosage`` (what BASELINE.json's configs exercise);
                  "literal": the condition table of the reference overrides every observed species'
                  initial value with 0.0 and never carries the dose (:369-376)
    """

    def __init__(self, model: str, parameters: Sequence[str], groups: Sequence[GroupData],
                 timepoints: Sequence[float], observables: Sequence[str], sigma=1.0,
                 priors: Optional[dict[str, tuple[float, float]]] = None, noise_model: int = runtime.NOISE_NORMAL,
                 mode: str = "intended", dosage: Optional[dict[str, float]] = None, device: int = 0,
                 rtol: float = 1e-8, atol: float = 1e-11, method: int = runtime.METHOD_AUTO,
                 prior_constant: bool = True):
        if mode not in ("intended", "literal"):
            raise ValueError("mode must be 'intended' or 'literal'")
        self.model = model
        self.parameters = list(parameters)
        self.groups = list(groups)
        self.n_groups = len(self.groups)
        self.observables = [o.strip("[]") for o in observables]
        self.timepoints = np.asarray(timepoints, dtype=np.float64)
        self.noise_model = noise_model
        self.x_names = [f"{par}_{g.id}" for par in self.parameters for g in self.groups]
        self.dim = len(self.x_names)
        self.lb = np.zeros(self.dim)
        self.ub = np.full(self.dim, 1e8)
        self.prior_constant = prior_constant
        self.prior_loc = np.full(self.dim, np.nan)
        self.prior_scale = np.full(self.dim, np.nan)
        for name, (loc, scale) in (priors or {}).items():
            i = self.x_names.index(name)
            self.prior_loc[i], self.prior_scale[i] = loc, scale
        self._has_prior = np.isfinite(self.prior_loc)

        pb = runtime.Problem(model, device=device)
        pb.set_solver(method=method, rtol=rtol, atol=atol)
        if mode == "literal":
            for o in self.observables:
                pb.set_value(o, 0.0)
        elif dosage:
            for k, v in dosage.items():
                pb.set_value(k, v)
        pb.set_columns(self.parameters)
        pb.set_timepoints(self.timepoints)
        pb.set_observables(self.observables)
        stats = np.stack([sufficient_statistics(g.y, noise_model) for g in self.groups], axis=-1)  # [3,T,K,G]
        pb.set_data(stats, sigma, noise_model)
        self.problem = pb
        # gradient rows come back in the model's compiled sensitivity order
        self._sens_row = [pb.sens_ids.index(p) if p in pb.sens_ids else -1 for p in self.parameters]
        self.n_evals = 0

    # ---- batched entry ---------------------------------------------------------------------------
    def logp_and_grad(self, X: np.ndarray, grad: bool = True) -> tuple[np.ndarray, Optional[np.ndarray]]:
        """X [B, dim] -> (log posterior [B], d log posterior / dx [B, dim] or None)."""
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        B, G, NP = X.shape[0], self.n_groups, len(self.parameters)
        if X.shape[1] != self.dim:
            raise ValueError(f"x must have {self.dim} entries ({self.x_names}), got {X.shape[1]}")
        if grad and min(self._sens_row) < 0:
            raise runtime.ParvarB200Error(
                f"gradient requested but {self.parameters} are not all compiled sensitivity parameters "
                f"{self.problem.sens_ids} of model {self.model}")
        # trajectory b*G + g carries theta_g of x_b:  P[p, b*G + g] = X[b, p*G + g]
        P = np.ascontiguousarray(X.reshape(B, NP, G).transpose(1, 0, 2).reshape(NP, B * G))
        lp_t, g_t, lp_x, n_failed = self.problem.logp_host(P, seg_len=G, grad=grad)
        self.n_evals += B
        logp = lp_x.copy()
        dx = None
        if grad:
            g_t = g_t[self._sens_row].reshape(NP, B, G)                 # [p, b, g]
            dx = np.ascontiguousarray(g_t.transpose(1, 0, 2).reshape(B, self.dim))
        if self._has_prior.any():
            z = (X[:, self._has_prior] - self.prior_loc[self._has_prior]) / self.prior_scale[self._has_prior]
            logp = logp - 0.5 * np.sum(z * z, axis=1)
            if self.prior_constant:
                logp = logp - 0.5 * np.sum(LOG_2PI + 2.0 * np.log(self.prior_scale[self._has_prior]))
            if grad:
                dx[:, self._has_prior] -= z / self.prior_scale[self._has_prior]
        if n_failed:
            bad = ~np.isfinite(lp_x)
            logp[bad] = -np.inf
            if grad:
                dx[bad] = np.nan
        return logp, dx

    # ---- pyPESTO-style single-point interface ----------------------------------------------------------
    def __call__(self, x, sensi_orders: tuple[int, ...] = (0,), mode: str = "mode_fun", return_dict: bool = False):
        if mode != "mode_fun":
            raise ValueError("only mode_fun is supported (residual mode is not part of the reference's path)")
        want_grad = 1 in sensi_orders
        logp, dx = self.logp_and_grad(np.asarray(x, dtype=np.float64)[None, :], grad=want_grad)
        fval = float(-logp[0])
        g = -dx[0] if want_grad else None
        if return_dict:
            out = {}
            if 0 in sensi_orders:
                out["fval"] = fval
            if want_grad:
                out["grad"] = g
            return out
        if 0 in sensi_orders and want_grad:
            return fval, g
        return g if want_grad else fval

    def fun(self, x) -> float:
        return self(x, (0,))

    def grad(self, x) -> np.ndarray:
        return self(x, (1,))

    def fun_and_grad(self, x) -> tuple[float, np.ndarray]:
        return self(x, (0, 1))


def make_problem(*args, **kwargs):
    """-> (fun, grad, lb, ub, x_names): the plain callables pyPESTO's ``Objective(fun=, grad=)`` takes."""
    obj = PetabObjective(*args, **kwargs)
    return obj.fun, obj.grad, obj.lb, obj.ub, obj.x_names
