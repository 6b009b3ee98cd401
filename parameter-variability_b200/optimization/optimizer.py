"""Batched multistart trust-region optimiser.

The reference runs ``pypesto.optimize.minimize`` with ``FidesOptimizer`` (maxiter 1e4, fatol = frtol = 1e-12), 100
uniformly drawn start points per problem, one after the other (src/parvar/optimization/petab_optimization.py:99-156,
settings :294-306); every fides iteration is one AMICI call returning fval, gradient and the Fisher information matrix
as Hessian approximation.  Here all starts of all problems advance together: one launch per iteration returns
(fval, grad, FIM) for every current iterate (``PetabObjectiveBatch.evaluate(order=2)``).

The step is a damped Gauss-Newton / Levenberg-Marquardt trust-region step with the FIM (+ prior precision) as model
Hessian, projected onto the box [lb, ub], with the classical gain-ratio update of the damping - the same model
(quadratic with FIM Hessian, box constraints) fides' default 2D-subspace trust-region method minimises; it is not a
line-by-line restatement of fides 0.8.0 (not in /root/reference).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np


@dataclass
class OptimizeResult:
    x: np.ndarray          # [Q, S, dim] final iterates, starts sorted by fval per problem
    fval: np.ndarray       # [Q, S]
    grad: np.ndarray       # [Q, S, dim]
    n_iter: np.ndarray     # [Q, S]
    converged: np.ndarray  # [Q, S] bool
    x0: np.ndarray         # [Q, S, dim]
    n_evals: int


def uniform_startpoints(n_starts: int, lb: np.ndarray, ub: np.ndarray, Q: int = 1, seed: int = 1234) -> np.ndarray:
    """``pypesto.startpoint.uniform`` (petab_optimization.py:120-121): uniform in [lb, ub]."""
    rs = np.random.RandomState(seed)
    return lb + (ub - lb) * rs.uniform(size=(Q, n_starts, len(lb)))


def minimize(evaluate: Callable[[np.ndarray], dict], x0: np.ndarray, lb: np.ndarray, ub: np.ndarray, maxiter: int = 200,
             fatol: float = 1e-12, frtol: float = 1e-12, gtol: float = 1e-8, lam0: float = 1e-2) -> OptimizeResult:
    """``evaluate(X [Q, S, dim]) -> {"fval", "grad", "hess"}`` (hess [Q, S, dim, dim] positive semi-definite)."""
    x = np.clip(np.asarray(x0, dtype=np.float64), lb, ub)
    Q, S, dim = x.shape
    out = evaluate(x)
    n_evals = Q * S
    f, g, H = out["fval"], out["grad"], out["hess"]
    lam = np.full((Q, S), lam0)
    active = np.isfinite(f)
    n_iter = np.zeros((Q, S), dtype=int)
    converged = np.zeros((Q, S), dtype=bool)
    eye = np.eye(dim)
    for it in range(maxiter):
        if not active.any():
            break
        Hs = np.where(np.isfinite(H), H, 0.0)
        gs = np.where(np.isfinite(g), g, 0.0)
        D = np.maximum(np.einsum("qsii->qsi", Hs), 1e-12)
        A = Hs + lam[..., None, None] * (D[..., None] * eye)
        try:
            p = -np.linalg.solve(A, gs[..., None])[..., 0]
        except np.linalg.LinAlgError:
            p = -gs / (D * (1 + lam[..., None]))
        x_try = np.clip(x + p, lb, ub)
        p = x_try - x
        pred = -(np.einsum("qsi,qsi->qs", gs, p) + 0.5 * np.einsum("qsi,qsij,qsj->qs", p, Hs, p))
        o = evaluate(np.where(active[..., None], x_try, x))
        n_evals += int(active.sum())
        f_try = np.where(np.isfinite(o["fval"]), o["fval"], np.inf)
        rho = (f - f_try) / np.maximum(pred, 1e-300)
        take = active & (f_try < f) & (rho > 1e-4)
        # convergence tests on accepted steps (fides: fatol / frtol on the objective decrease, gtol on the gradient)
        df = np.where(take, f - f_try, np.inf)
        x = np.where(take[..., None], x_try, x)
        g = np.where(take[..., None], o["grad"], g)
        H = np.where(take[..., None, None], o["hess"], H)
        f = np.where(take, f_try, f)
        lam = np.where(take, np.maximum(lam * np.where(rho > 0.75, 1.0 / 3.0, 1.0), 1e-12), np.minimum(lam * 4.0, 1e12))
        n_iter += active
        # projected gradient norm
        gp = np.where(((x <= lb) & (g > 0)) | ((x >= ub) & (g < 0)), 0.0, g)
        done = active & ((df < fatol + frtol * np.abs(f)) | (np.max(np.abs(gp), axis=2) < gtol) | (lam >= 1e12)
                         | (np.max(np.abs(p), axis=2) < 1e-14 * (1 + np.max(np.abs(x), axis=2))))
        converged |= done & np.isfinite(f)
        active &= ~done
    order = np.argsort(np.where(np.isfinite(f), f, np.inf), axis=1)
    take = lambda a: np.take_along_axis(a, order[(...,) + (None,) * (a.ndim - 2)], axis=1)  # noqa: E731
    return OptimizeResult(x=take(x), fval=take(f), grad=take(g), n_iter=take(n_iter), converged=take(converged),
                          x0=take(np.clip(np.asarray(x0, dtype=np.float64), lb, ub)), n_evals=n_evals)


def minimize_device(objective, x0: np.ndarray, maxiter: int = 200, fatol: float = 1e-12, frtol: float = 1e-12, gtol: float = 1e-8,
                    poll_every: int = 4) -> OptimizeResult:
    """``minimize`` with the whole loop on the GPU (``pv_optimizer_create`` / ``run`` / ``result``): iterate, gradient, FIM,
    damping and convergence flags of every start of every problem stay in HBM; one iteration is step kernel -> batched
    objective with Fisher information -> update kernel.  ``objective`` is a ``PetabObjectiveBatch`` (problem, priors and
    bounds are taken from it).  Same steps, same accept and stopping rules as the numpy statement above."""
    import ctypes
    from .. import runtime
    lib = runtime.load_library() if objective.problem._lib is None else objective.problem._lib
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    Q, S, dim = x0.shape
    if Q != objective.Q or dim != objective.dim:
        raise ValueError(f"x0 must be [{objective.Q}, S, {objective.dim}]")
    lb = np.ascontiguousarray(np.broadcast_to(np.asarray(objective.lb, dtype=np.float64), (dim,)))
    ub = np.ascontiguousarray(np.broadcast_to(np.asarray(objective.ub, dtype=np.float64), (dim,)))
    ploc = np.ascontiguousarray(objective.prior_loc, dtype=np.float64)
    pscale = np.ascontiguousarray(np.where(np.isfinite(objective.prior_loc), objective.prior_scale, 1.0), dtype=np.float64)
    ptr = lambda a: a.ctypes.data   # noqa: E731
    h = lib.pv_optimizer_create(objective.problem._h, Q, objective.G, S, ptr(lb), ptr(ub), ptr(ploc), ptr(pscale), ptr(x0),
                                float(fatol), float(frtol), float(gtol))
    if not h:
        raise runtime.ParvarB200Error(f"pv_optimizer_create failed: {runtime.last_error(lib)}")
    try:
        done = ctypes.c_int(0)
        runtime._check(lib.pv_optimizer_run(h, int(maxiter), int(poll_every), ctypes.addressof(done)), lib)
        x, f, g = np.empty((Q, S, dim)), np.empty((Q, S)), np.empty((Q, S, dim))
        n_iter, conv = np.empty((Q, S), dtype=np.int32), np.empty((Q, S), dtype=np.uint8)
        ne = ctypes.c_int64(0)
        runtime._check(lib.pv_optimizer_result(h, ptr(x), ptr(f), ptr(g), ptr(n_iter), ptr(conv), ctypes.addressof(ne)), lib)
    finally:
        lib.pv_optimizer_destroy(h)
    objective.n_evals += int(ne.value)
    order = np.argsort(np.where(np.isfinite(f), f, np.inf), axis=1, kind="stable")
    take = lambda a: np.take_along_axis(a, order[(...,) + (None,) * (a.ndim - 2)], axis=1)  # noqa: E731
    return OptimizeResult(x=take(x), fval=take(f), grad=take(g), n_iter=take(n_iter.astype(int)), converged=take(conv.astype(bool)),
                          x0=take(np.clip(x0, lb, ub)), n_evals=int(ne.value))
