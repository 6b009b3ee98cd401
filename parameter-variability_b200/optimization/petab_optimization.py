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


def model_with_sensitivities(model: str, parameters: Sequence[str]) -> str:
    """The compiled model to use for gradients with respect to ``parameters``: the registry build when they are among its
    sensitivity parameters, otherwise a run-time build of the same SBML file with exactly this set (generated, compiled
    and cached on first use - AMICI likewise builds its model for whatever the PEtab table estimates,
    petab_optimization.py:43-45)."""
    from .. import MODELS, codegen
    if model not in codegen.BUILTIN or set(parameters) <= set(codegen.BUILTIN[model]):
        return model
    from ..experiments.petab_factory import resolve_model
    return resolve_model(MODELS[model], sens=list(parameters))


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
                 prior_constant: bool = True, max_steps: int = 10000, t0: float = 0.0):
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

        pb = runtime.Problem(model_with_sensitivities(model, self.parameters), device=device)
        # max_steps = 10000 is AMICI's default step budget: a parameter vector that needs more (start points and tempered
        # chains roam the PEtab bounds [0, 1e8]) is a failed evaluation (fval = inf), not a minutes-long integration
        pb.set_solver(method=method, rtol=rtol, atol=atol, max_steps=max_steps)
        if mode == "literal":
            for o in self.observables:
                pb.set_value(o, 0.0)
        elif dosage:
            for k, v in dosage.items():
                pb.set_value(k, v)
        pb.set_columns(self.parameters)
        # PEtab / AMICI: the initial values hold at t = 0 even when the first measurement is later
        pb.set_timepoints(self.timepoints, t0=min(float(t0), float(self.timepoints[0])))
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
        # a failed integration (NaN from the kernel) and a non-finite likelihood of a finished one (log-normal noise with
        # f <= 0) are both failed evaluations: fval = inf, grad = nan - whatever n_failed says
        bad = ~np.isfinite(lp_x)
        if bad.any():
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


class PetabObjectiveBatch:
    """Q independent PEtab problems of one model (same parameters, time grid, observables; different measurements
    and priors) evaluated together: ``X [Q, B, dim] -> loglik [Q, B], logprior [Q, B], grad [Q, B, dim], fim [Q, B, dim, dim]``.

    This is the reference's hyper-parameter sweep (3 x 2160 problems, ``factories/__init__.py:12-20``) turned from a
    process pool over problems (``petab_optimization.py:381-391``) into ONE launch per evaluation: the problem index
    is just another data set (trajectory ((b * Q + q) * G + g) reads data set q * G + g).
    ``PetabObjective`` is the Q = 1 view with pyPESTO's calling conventions.
    """

    def __init__(self, model: str, parameters: Sequence[str], problems: Sequence[Sequence[GroupData]],
                 timepoints: Sequence[float], observables: Sequence[str], sigma=1.0,
                 priors: Optional[Sequence[dict[str, tuple[float, float]]]] = None,
                 noise_model: int = runtime.NOISE_NORMAL, mode: str = "intended",
                 dosage: Optional[dict[str, float]] = None, device: int = 0, rtol: float = 1e-8, atol: float = 1e-11,
                 method: int = runtime.METHOD_AUTO, t0: float = 0.0, max_steps: int = 10000,
                 initial_values: Optional[dict[str, float]] = None):
        self.model = model
        self.parameters = list(parameters)
        self.Q = len(problems)
        self.group_ids = [g.id for g in problems[0]]
        self.G = len(self.group_ids)
        for pr in problems:
            if [g.id for g in pr] != self.group_ids:
                raise ValueError("all problems must have the same groups")
        self.observables = [o.strip("[]") for o in observables]
        self.timepoints = np.asarray(timepoints, dtype=np.float64)
        self.x_names = [f"{par}_{g}" for par in self.parameters for g in self.group_ids]
        self.dim = len(self.x_names)
        self.lb = np.zeros(self.dim)
        self.ub = np.full(self.dim, 1e8)
        self.prior_loc = np.full((self.Q, self.dim), np.nan)
        self.prior_scale = np.full((self.Q, self.dim), np.nan)
        for q, pri in enumerate(priors or []):
            for name, (loc, scale) in pri.items():
                i = self.x_names.index(name)
                self.prior_loc[q, i], self.prior_scale[q, i] = loc, scale
        self.has_prior = np.isfinite(self.prior_loc)

        pb = runtime.Problem(model_with_sensitivities(model, self.parameters), device=device)
        pb.set_solver(method=method, rtol=rtol, atol=atol, max_steps=max_steps)      # AMICI's default budget, see PetabObjective
        if mode == "literal":
            # the condition table's species columns when given (read from the files), else what the reference always
            # writes: every observed species 0.0 (quirk Q4)
            for k, v in (initial_values if initial_values is not None else {o: 0.0 for o in self.observables}).items():
                pb.set_value(k, v)
        elif mode == "intended":
            for k, v in (dosage or {}).items():
                pb.set_value(k, v)
        else:
            raise ValueError("mode must be 'intended' or 'literal'")
        pb.set_columns(self.parameters)
        pb.set_timepoints(self.timepoints, t0=min(t0, float(self.timepoints[0])))
        pb.set_observables(self.observables)
        stats = np.stack([sufficient_statistics(g.y, noise_model) for pr in problems for g in pr], axis=-1)  # [3,T,K,Q*G]
        pb.set_data(stats, sigma, noise_model)
        self.problem = pb
        self._sens_row = [pb.sens_ids.index(p) if p in pb.sens_ids else -1 for p in self.parameters]
        self.n_evals = 0

    def log_prior(self, X: np.ndarray, grad: bool = False):
        """parameterScaleNormal priors on the lin scale (petab_factory.py:443-449), constant included."""
        z = np.where(self.has_prior[:, None, :], (X - self.prior_loc[:, None, :]) / self.prior_scale[:, None, :], 0.0)
        const = np.where(self.has_prior, LOG_2PI + 2.0 * np.log(np.where(self.has_prior, self.prior_scale, 1.0)), 0.0)
        lp = -0.5 * np.sum(z * z, axis=2) - 0.5 * np.sum(const, axis=1)[:, None]
        if not grad:
            return lp, None, None
        inv = np.where(self.has_prior, 1.0 / np.where(self.has_prior, self.prior_scale, 1.0), 0.0)
        return lp, -z * inv[:, None, :], (inv * inv)[:, None, :] * np.ones_like(X)

    def evaluate(self, X: np.ndarray, order: int = 0):
        """X [Q, B, dim].  order 0: (llh, lprior); 1: + (dllh, dlprior); 2: + (fim [Q,B,dim,dim], prior precision diag)."""
        X = np.asarray(X, dtype=np.float64)
        Q, B, dim = X.shape
        if Q != self.Q or dim != self.dim:
            raise ValueError(f"X must be [{self.Q}, B, {self.dim}]")
        G, NP = self.G, len(self.parameters)
        if order > 0 and min(self._sens_row) < 0:
            raise runtime.ParvarB200Error(f"{self.parameters} are not all compiled sensitivity parameters of {self.model}")
        # trajectory ((b*Q + q)*G + g): theta_g of X[q, b]  ->  P[p, (b*Q+q)*G + g] = X[q, b, p*G + g]
        P = np.ascontiguousarray(X.reshape(Q, B, NP, G).transpose(2, 1, 0, 3).reshape(NP, B * Q * G))
        self.n_evals += Q * B
        out = {}
        if order == 0:
            lp_t, _, seg, nf = self.problem.logp_host(P, seg_len=G, grad=False)
        elif order == 1:
            lp_t, g_t, seg, nf = self.problem.logp_host(P, seg_len=G, grad=True)
        else:
            lp_t, g_t, f_t, seg, nf = self.problem.logp_fim_host(P, seg_len=G)
        llh = seg.reshape(B, Q).T.copy()
        bad = ~np.isfinite(llh)
        llh[bad] = -np.inf
        lprior, dprior, hprior = self.log_prior(X, grad=order > 0)
        out["llh"], out["lprior"] = llh, lprior
        if order > 0:
            g_t = g_t[self._sens_row].reshape(NP, B, Q, G)                       # [p, b, q, g]
            dllh = np.ascontiguousarray(g_t.transpose(2, 1, 0, 3).reshape(Q, B, dim))
            dllh[bad] = np.nan
            out["dllh"], out["dlprior"] = dllh, dprior
        if order > 1:
            ns = self.problem.n_sens
            tri = [(s, r) for s in range(ns) for r in range(s, ns)]
            F = np.zeros((Q, B, dim, dim))
            f_t = f_t.reshape(len(tri), B, Q, G)
            for k, (s, r) in enumerate(tri):
                # sens index -> position among self.parameters
                if s not in self._sens_row or r not in self._sens_row:
                    continue
                ps, pr_ = self._sens_row.index(s), self._sens_row.index(r)
                for g in range(G):
                    v = f_t[k, :, :, g].T                                          # [Q, B]
                    F[:, :, ps * G + g, pr_ * G + g] = v
                    F[:, :, pr_ * G + g, ps * G + g] = v
            out["fim"], out["prior_precision"] = F, hprior
        return out


class HierarchicalObjective:
    """Population (hierarchical) form of BASELINE.json configs 3-4: every individual i has its own
    theta_i = (theta_i1 .. theta_iP) and its own measurements; theta_ip ~ LogNormal(mu_p, omega_p); hyper-priors
    mu_p ~ Normal(m_p, s_p), omega_p ~ HalfNormal(h_p).  (Build-defined: the reference has no hierarchy)

        log p(theta, mu, omega | y) = sum_i loglik_i(theta_i) + sum_i,p log LogNormal(theta_ip; mu_p, omega_p)
                                      + sum_p [log N(mu_p; m_p, s_p) + log HalfNormal(omega_p; h_p)]

    The per-individual likelihoods and their gradients come from ONE kernel launch over n_chains x n_individuals
    trajectories (``Problem.logp_grad`` with ``seg_len = n_individuals``); the population terms are O(n) host algebra.
    With several ranks the individuals are sharded and ``distributed.ShardedLikelihood`` all-reduces
    ``[n_chains, 1 + 2P]`` (logp and the hyper-parameter gradient).
    """

    def __init__(self, problem: "runtime.Problem", n_individuals: int, mu_prior: Sequence[tuple[float, float]],
                 omega_prior: Sequence[float], first_individual: int = 0, n_local: Optional[int] = None):
        self.problem = problem
        self.n_ind = n_individuals
        self.first = first_individual
        self.n_local = n_individuals if n_local is None else n_local
        self.P = len(problem.columns)
        self.mu_prior = np.asarray(mu_prior, dtype=np.float64)       # [P, 2]
        self.omega_prior = np.asarray(omega_prior, dtype=np.float64)  # [P]
        self._sens_row = [problem.sens_ids.index(c) for c in problem.columns]

    def local_terms(self, theta: np.ndarray, mu: np.ndarray, omega: np.ndarray):
        """theta [C, n_local, P], mu / omega [C, P] -> (logp_local [C], dtheta [C, n_local, P], dmu [C, P], domega [C, P])
        for this rank's individuals (likelihood + population density); hyper-priors are added once by ``finish``."""
        C, L, P = theta.shape
        # parameter-major [P, C, L] throughout: it is the kernel's own layout, so nothing is transposed around the launch
        thp = np.ascontiguousarray(theta.transpose(2, 0, 1))
        lp, g, seg, nf = self.problem.logp_host(thp.reshape(P, C * L), seg_len=L, grad=True)
        self.n_failed = nf
        g = g[self._sens_row].reshape(P, C, L)
        iw = (1.0 / omega).T[:, :, None]                               # [P, C, 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            lt = np.log(thp)
        z = lt - mu.T[:, :, None]
        z *= iw
        s_lt, s_z, s_zz = lt.sum(axis=2), z.sum(axis=2), np.einsum("pcl,pcl->pc", z, z)     # [P, C]
        pop = -s_lt.sum(axis=0) - L * np.log(omega).sum(axis=1) - 0.5 * LOG_2PI * L * P - 0.5 * s_zz.sum(axis=0)
        z *= iw
        z += 1.0
        with np.errstate(divide="ignore", invalid="ignore"):   # theta <= 0 (a leapfrog overshoot): NaN gradient, logp = NaN
            z /= thp
        g -= z                                                         # d/dtheta of the population density: -(1 + z/omega)/theta
        dmu = np.ascontiguousarray((s_z * iw[:, :, 0]).T)            # C-contiguous results: callers pack them into
        domega = np.ascontiguousarray(((s_zz - L) * iw[:, :, 0]).T)  # all-reduce buffers
        return seg + pop, np.ascontiguousarray(g.transpose(1, 2, 0)), dmu, domega

    def finish(self, logp: np.ndarray, dmu: np.ndarray, domega: np.ndarray, mu: np.ndarray, omega: np.ndarray):
        """Add the hyper-priors (after the cross-rank reduction)."""
        m, s = self.mu_prior[:, 0], self.mu_prior[:, 1]
        h = self.omega_prior
        zm = (mu - m) / s
        logp = logp + np.sum(-0.5 * (LOG_2PI + 2 * np.log(s)) - 0.5 * zm * zm, axis=1)
        logp = logp + np.sum(0.5 * np.log(2.0 / np.pi) - np.log(h) - 0.5 * (omega / h) ** 2, axis=1)
        return logp, dmu - zm / s, domega - omega / (h * h)

    def __call__(self, theta: np.ndarray, mu: np.ndarray, omega: np.ndarray):
        lp, dth, dmu, dom = self.local_terms(theta, mu, omega)
        lp, dmu, dom = self.finish(lp, dmu, dom, mu, omega)
        return lp, dth, dmu, dom


class HierarchicalObjectiveDevice:
    """``HierarchicalObjective`` with nothing on the host: theta, mu, omega are CUDA tensors, the population terms and
    the per-chain sums are an epilogue kernel behind the sensitivity kernel (``pv_hier_logp_grad``), the buffer
    ``[C, 1 + 2P]`` it writes is all-reduced in place by NCCL on the same stream when the individuals of a chain are
    sharded over ranks, and the hyper-priors are added by ``pv_hier_finish``.  An evaluation
    is 3 launches + at most one collective, queued without a host synchronisation; the results stay on the device for
    the sampler's next leapfrog step.

    Data set i of ``problem`` must be this rank's i-th individual (``n_data == n_local``).
    """

    def __init__(self, problem: "runtime.Problem", n_chains: int, n_individuals: int,
                 mu_prior: Sequence[tuple[float, float]], omega_prior: Sequence[float], first_individual: int = 0,
                 n_local: Optional[int] = None, group=None, world: int = 1):
        import torch
        self._torch = torch
        self.problem = problem
        self.C = int(n_chains)
        self.n_ind = int(n_individuals)
        self.first = int(first_individual)
        self.n_local = self.n_ind if n_local is None else int(n_local)
        self.P = len(problem.columns)
        if problem.n_data not in (0, self.n_local) and self.n_local > 0:
            raise ValueError(f"problem holds {problem.n_data} data sets, expected one per local individual ({self.n_local})")
        mp = np.asarray(mu_prior, dtype=np.float64).reshape(self.P, 2)
        self.mu_prior_mean, self.mu_prior_sd = np.ascontiguousarray(mp[:, 0]), np.ascontiguousarray(mp[:, 1])
        self.omega_prior = np.ascontiguousarray(omega_prior, dtype=np.float64)
        self.group, self.world = group, int(world)
        dev = torch.device("cuda", problem.device)
        B = self.C * self.n_local
        self.dtheta = torch.empty((self.P, self.C, self.n_local), dtype=torch.float64, device=dev)
        self.red = torch.empty((self.C, 1 + 2 * self.P), dtype=torch.float64, device=dev)
        self.status = torch.zeros(max(B, 1), dtype=torch.int32, device=dev)

    def __call__(self, theta, mu, omega):
        """theta [P, C, n_local] (parameter-major: the kernel's own layout), mu / omega [C, P], CUDA float64 tensors.
        -> (red [C, 1 + 2P] = (log posterior, d/d mu, d/d omega) per chain, dtheta [P, C, n_local]); both are this
        object's buffers, valid until the next call, ordered on torch's current stream."""
        torch = self._torch
        stream = torch.cuda.current_stream().cuda_stream
        self.problem.hier_logp_grad(self.C, self.n_local, theta, mu, omega, self.dtheta, self.red, status=self.status,
                                    stream=stream)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.red, op=dist.ReduceOp.SUM, group=self.group)     # NCCL, ordered after the epilogue kernel
        runtime.hier_finish(self.C, self.P, self.red, mu, omega, self.mu_prior_mean, self.mu_prior_sd, self.omega_prior,
                            stream=stream)
        return self.red, self.dtheta
