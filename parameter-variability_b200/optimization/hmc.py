"""Batched Hamiltonian Monte Carlo for the gradient path (BASELINE.json north_star: "the log-likelihood and its
gradient that the MCMC/NUTS sampler calls at every step").

All chains advance in lock step: one leapfrog step of every chain = ONE batched ``logp_and_grad`` evaluation, which for
the hierarchical objective is one kernel launch over n_chains x n_individuals trajectories with forward sensitivities
(BASELINE config 3: "time 10^3 consecutive evaluations (leapfrog-like usage)").

Static-length HMC (Neal 2011) with jittered trajectory length, dual-averaging step-size adaptation to a target
acceptance rate (Hoffman & Gelman 2014, section 3.2) and a diagonal mass matrix estimated during warm-up (Stan's
windowed scheme reduced to one window).  The no-U-turn criterion itself is not implemented: its tree depth differs per
chain and would serialise the batch; fixed-length trajectories keep every launch full.
Positive parameters are sampled on the log scale by the caller (see ``tests``), so the sampler itself is unconstrained.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np


@dataclass
class HMCResult:
    samples: np.ndarray        # [n_chains, n_samples, dim]
    logp: np.ndarray           # [n_chains, n_samples]
    accept_rate: np.ndarray    # [n_chains] mean acceptance probability after warm-up
    step_size: np.ndarray      # [n_chains] adapted step size
    inv_mass: np.ndarray       # [n_chains, dim] adapted diagonal inverse mass
    n_grad_evals: int          # batched gradient evaluations (= kernel launches of the objective)
    divergences: np.ndarray    # [n_chains]


def hmc(logp_and_grad: Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]], x0: np.ndarray, n_samples: int = 500,
        n_warmup: int = 500, n_leapfrog: int = 16, step_size: float = 0.05, target_accept: float = 0.8, seed: int = 1234,
        adapt_mass: bool = True, max_energy_error: float = 1000.0) -> HMCResult:
    """``logp_and_grad(X [C, dim]) -> (logp [C], grad [C, dim])``; ``x0`` [C, dim]."""
    rs = np.random.RandomState(seed)
    x = np.array(x0, dtype=np.float64)
    C, dim = x.shape
    lp, g = logp_and_grad(x)
    n_evals = 1
    eps = np.full(C, step_size)
    inv_mass = np.ones((C, dim))
    # dual averaging state (Hoffman & Gelman, Algorithm 5)
    mu = np.log(10.0 * eps)
    log_eps_bar = np.zeros(C)
    h_bar = np.zeros(C)
    gamma, t0, kappa = 0.05, 10.0, 0.75
    # running moments for the mass matrix (Welford), second half of warm-up
    w_n, w_mean, w_m2 = 0, np.zeros((C, dim)), np.zeros((C, dim))
    samples = np.empty((C, n_samples, dim))
    lps = np.empty((C, n_samples))
    acc_sum = np.zeros(C)
    div = np.zeros(C, dtype=int)
    for it in range(n_warmup + n_samples):
        warm = it < n_warmup
        p = rs.normal(size=(C, dim)) / np.sqrt(inv_mass)
        h0 = -lp + 0.5 * np.sum(p * p * inv_mass, axis=1)
        xq, pq, gq, lpq = x.copy(), p.copy(), g.copy(), lp.copy()
        L = max(1, int(round(n_leapfrog * rs.uniform(0.8, 1.2))))
        ok = np.ones(C, dtype=bool)
        for _ in range(L):
            pq = pq + 0.5 * eps[:, None] * gq
            xq = xq + eps[:, None] * inv_mass * pq
            lpq, gq = logp_and_grad(xq)
            n_evals += 1
            bad = ~np.isfinite(lpq) | ~np.all(np.isfinite(gq), axis=1)
            ok &= ~bad
            gq = np.where(bad[:, None], 0.0, gq)
            pq = pq + 0.5 * eps[:, None] * gq
        h1 = -lpq + 0.5 * np.sum(pq * pq * inv_mass, axis=1)
        dh = np.where(ok, h0 - h1, -np.inf)
        diverged = ~ok | (dh < -max_energy_error)
        a = np.where(diverged, 0.0, np.minimum(1.0, np.exp(np.minimum(dh, 0.0))))
        accept = rs.uniform(size=C) < a
        x = np.where(accept[:, None], xq, x)
        g = np.where(accept[:, None], gq, g)
        lp = np.where(accept, lpq, lp)
        if warm:
            m = it + 1
            h_bar = (1 - 1 / (m + t0)) * h_bar + (target_accept - a) / (m + t0)
            log_eps = mu - np.sqrt(m) / gamma * h_bar
            eta = m ** (-kappa)
            log_eps_bar = eta * log_eps + (1 - eta) * log_eps_bar
            eps = np.exp(log_eps)
            if adapt_mass and it >= n_warmup // 2:
                w_n += 1
                d = x - w_mean
                w_mean += d / w_n
                w_m2 += d * (x - w_mean)
            if it == n_warmup - 1:
                eps = np.exp(log_eps_bar)
                if adapt_mass and w_n > 10:
                    var = w_m2 / (w_n - 1)
                    inv_mass = (w_n / (w_n + 5.0)) * var + 1e-3 * (5.0 / (w_n + 5.0))      # Stan's shrinkage
        else:
            k = it - n_warmup
            samples[:, k], lps[:, k] = x, lp
            acc_sum += a
            div += diverged
    return HMCResult(samples=samples, logp=lps, accept_rate=acc_sum / max(n_samples, 1), step_size=eps, inv_mass=inv_mass,
                     n_grad_evals=n_evals, divergences=div)
