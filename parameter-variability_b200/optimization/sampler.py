"""Batched adaptive-Metropolis + adaptive parallel tempering.

The reference samples each PEtab problem with
``pypesto.sample.AdaptiveParallelTemperingSampler(internal_sampler=AdaptiveMetropolisSampler(), n_chains=4)`` for
5000 samples (src/parvar/optimization/petab_optimization.py:158-173, settings :294-306): 4 x 5000 sequential Python ->
AMICI objective calls per problem, problems spread over a process pool.  Here every iteration evaluates the proposals
of ALL chains of ALL problems in one kernel launch (``PetabObjectiveBatch.evaluate``); the bookkeeping around it is
compiled host code over [problems, chains] (``pv_sampler_propose`` / ``pv_sampler_step`` of the C ABI; the numpy
statement of the same loop is kept as ``native=False``: same random stream, same chain up to round-off).

Algorithms (restated from the publications pyPESTO implements; pyPESTO 0.6.0 is not in /root/reference):
  * adaptive Metropolis: Haario, Saksman, Tamminen (2001) with the decaying update of the running mean / covariance
    and of the proposal scale towards the 0.234 acceptance rate (pyPESTO ``AdaptiveMetropolisSampler``:
    decay_constant 0.51, reg_factor 1e-6, target_acceptance_rate 0.234, threshold_sample 1);
  * parallel tempering with swaps between neighbouring temperatures, likelihood tempered (posterior_beta
    ~ L^beta * prior), initial ladder beta_k = 1 / T_k with near-exponentially spaced temperatures (exponent 4,
    max_temp 5e4), ladder adapted as in Vousden, Farr, Mandel (2016) (nu = 1e3, eta = 100).
Proposals outside the bounds are rejected (pyPESTO: objective = inf outside [lb, ub]).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np


def near_exponential_decay_betas(n_chains: int, exponent: float = 4.0, max_temp: float = 5e4) -> np.ndarray:
    if n_chains == 1:
        return np.array([1.0])
    temps = 1.0 + np.linspace(0.0, 1.0, n_chains) ** exponent * (max_temp - 1.0) if np.isfinite(max_temp) else None
    if temps is None:
        temps = np.linspace(1.0, np.inf, n_chains)
    return 1.0 / temps


def _positive_definite(a: np.ndarray) -> np.ndarray:
    """Batched test ``a`` [..., n, n] (symmetric) is strictly positive definite: an unrolled Cholesky factorisation,
    vectorised over the batch (n is the parameter dimension of a problem, 2..6 here), all pivots > 0."""
    n = a.shape[-1]
    low = np.zeros_like(a)
    ok = np.ones(a.shape[:-2], dtype=bool)
    for j in range(n):
        d = a[..., j, j] - np.einsum("...k,...k->...", low[..., j, :j], low[..., j, :j])
        ok &= d > 0
        dj = np.sqrt(np.where(d > 0, d, 1.0))
        low[..., j, j] = dj
        if j + 1 < n:
            low[..., j + 1:, j] = (a[..., j + 1:, j] - np.einsum("...ik,...k->...i", low[..., j + 1:, :j], low[..., j, :j])) / dj[..., None]
    return ok


def _regularize(cov: np.ndarray, reg: float) -> np.ndarray:
    """Symmetrise and make positive definite.  pyPESTO adds reg_factor * I until the matrix is PD; the hot chains of a
    problem with bounds [0, 1e8] reach variances of 1e15, where an absolute 1e-6 no longer cures round-off, so the
    eigenvalues are floored relative to the largest one as well: lambda >= max(reg, 1e-12 lambda_max).

    This runs for every chain of every problem at every iteration, and a batched ``eigh`` was 2/3 of the sampler's host
    time.  Almost every matrix is fine, which a Cholesky test shows cheaply: cov - f I is positive definite for
    f = max(reg, 1e-12 trace) >= the floor above.  Only the matrices that fail it go through the eigen-decomposition."""
    cov = 0.5 * (cov + np.swapaxes(cov, -1, -2))
    n = cov.shape[-1]
    bound = np.maximum(reg, 1e-12 * np.abs(np.trace(cov, axis1=-2, axis2=-1)))
    ok = _positive_definite(cov - bound[..., None, None] * np.eye(n))
    if np.all(ok):
        return cov
    bad = ~ok
    w, V = np.linalg.eigh(cov[bad])
    w = np.maximum(w, np.maximum(reg, 1e-12 * np.abs(w[..., -1:])))
    cov[bad] = np.einsum("...ik,...k,...jk->...ij", V, w, V)
    return cov


@dataclass
class SampleResult:
    trace_x: np.ndarray            # [Q, n_chains, n_samples + 1, dim]
    trace_neglogpost: np.ndarray   # [Q, n_chains, n_samples + 1]
    trace_neglogprior: np.ndarray  # [Q, n_chains, n_samples + 1]
    betas: np.ndarray              # [Q, n_chains] final temperature ladder
    acceptance_rate: np.ndarray    # [Q, n_chains]
    swap_rate: np.ndarray          # [Q, n_chains - 1]
    n_evals: int
    burn_in: Optional[np.ndarray] = None
    effective_sample_size: Optional[np.ndarray] = None
    timing: Optional[dict] = None  # device loop: seconds spent drawing random numbers / queueing + waiting / copying out


def sample(evaluate: Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]], x0: np.ndarray, lb: np.ndarray, ub: np.ndarray,
           n_samples: int = 5000, n_chains: int = 4, seed: int = 1234, cov0: Optional[np.ndarray] = None,
           decay_constant: float = 0.51, target_acceptance_rate: float = 0.234, reg_factor: float = 1e-6,
           threshold_sample: int = 1, nu: float = 1e3, eta: float = 100.0, adapt_betas: bool = True,
           native: Optional[bool] = None) -> SampleResult:
    """``evaluate(X [Q, C, dim]) -> (llh [Q, C], lprior [Q, C])``; ``x0`` [Q, dim] (start of every chain, as pyPESTO
    starts all tempered chains at the best optimisation result).  ``native``: bookkeeping in the library's compiled host
    code (default) or in numpy."""
    if native is None or native:
        return _sample_native(evaluate, x0, lb, ub, n_samples, n_chains, seed, cov0, decay_constant, target_acceptance_rate,
                              reg_factor, threshold_sample, nu, eta, adapt_betas)
    rs = np.random.RandomState(seed)
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    Q, dim = x0.shape
    C = n_chains
    x = np.repeat(x0[:, None, :], C, axis=1)
    llh, lpr = evaluate(x)
    n_evals = Q * C
    betas = np.repeat(near_exponential_decay_betas(C)[None, :], Q, axis=0)
    cov = np.broadcast_to(np.eye(dim) if cov0 is None else np.asarray(cov0, dtype=np.float64), (Q, C, dim, dim)).copy()
    cov = _regularize(cov, reg_factor)
    mean_hist = x.copy()
    cov_hist = cov.copy()
    cov_scale = np.ones((Q, C))
    trace_x = np.empty((Q, C, n_samples + 1, dim))
    trace_lp = np.empty((Q, C, n_samples + 1))
    trace_pr = np.empty((Q, C, n_samples + 1))
    trace_x[:, :, 0], trace_lp[:, :, 0], trace_pr[:, :, 0] = x, -(llh + lpr), -lpr
    n_acc = np.zeros((Q, C))
    n_swap = np.zeros((Q, max(C - 1, 1)))
    n_swap_try = np.zeros((Q, max(C - 1, 1)))

    for it in range(1, n_samples + 1):
        # ---- Metropolis step for every chain of every problem -------------------------------------------------------
        Lc = np.linalg.cholesky(cov)
        x_new = x + np.einsum("qcij,qcj->qci", Lc, rs.normal(size=(Q, C, dim)))
        inside = np.all((x_new >= lb) & (x_new <= ub), axis=2)
        llh_new, lpr_new = evaluate(np.where(inside[..., None], x_new, x))
        n_evals += Q * C
        llh_new = np.where(inside, llh_new, -np.inf)
        with np.errstate(invalid="ignore"):
            log_acc = betas * (llh_new - llh) + (lpr_new - lpr)
        log_acc = np.where(np.isnan(log_acc), -np.inf, np.minimum(log_acc, 0.0))
        accept = np.log(rs.uniform(size=(Q, C))) < log_acc
        x = np.where(accept[..., None], x_new, x)
        llh = np.where(accept, llh_new, llh)
        lpr = np.where(accept, lpr_new, lpr)
        n_acc += accept
        # ---- adapt proposal (Haario et al.) ---------------------------------------------------------------------------
        if it > threshold_sample:
            rate = it ** (-decay_constant)
            mean_hist = (1 - rate) * mean_hist + rate * x
            dx = x - mean_hist
            cov_hist = (1 - rate) * cov_hist + rate * np.einsum("qci,qcj->qcij", dx, dx)
            cov_scale = cov_scale * np.exp((np.exp(log_acc) - target_acceptance_rate) / (it + 1) ** decay_constant)
            cov = _regularize(cov_scale[..., None, None] * cov_hist, reg_factor)
        # ---- swaps between neighbouring temperatures, hottest pair first ---------------------------------------------
        swapped = np.zeros((Q, max(C - 1, 1)), dtype=bool)
        for k in range(C - 1, 0, -1):
            dbeta = betas[:, k - 1] - betas[:, k]
            with np.errstate(invalid="ignore"):
                p_swap = dbeta * (llh[:, k] - llh[:, k - 1])
            p_swap = np.where(np.isnan(p_swap), -np.inf, p_swap)
            do = np.log(rs.uniform(size=Q)) < p_swap
            n_swap_try[:, k - 1] += 1
            n_swap[:, k - 1] += do
            swapped[:, k - 1] = do
            for arr in (x, llh, lpr):
                a, b = arr[:, k - 1].copy(), arr[:, k].copy()
                sel = do[:, None] if arr.ndim == 3 else do
                arr[:, k - 1] = np.where(sel, b, a)
                arr[:, k] = np.where(sel, a, b)
        # ---- adapt the ladder (Vousden et al. 2016) ---------------------------------------------------------------------
        if adapt_betas and C > 2:
            kappa = nu / (it + 1 + nu) / eta
            ds = kappa * (swapped[:, :-1].astype(float) - swapped[:, 1:].astype(float))
            with np.errstate(divide="ignore"):
                temps = 1.0 / betas
            dtemps = np.diff(temps[:, :-1], axis=1) * np.exp(ds[:, :C - 2])
            temps[:, 1:-1] = temps[:, :1] + np.cumsum(dtemps, axis=1)
            betas = 1.0 / temps
        trace_x[:, :, it], trace_lp[:, :, it], trace_pr[:, :, it] = x, -(llh + lpr), -lpr

    return SampleResult(trace_x=trace_x, trace_neglogpost=trace_lp, trace_neglogprior=trace_pr, betas=betas,
                        acceptance_rate=n_acc / n_samples, swap_rate=n_swap / np.maximum(n_swap_try, 1), n_evals=n_evals)


def _sample_native(evaluate, x0, lb, ub, n_samples, n_chains, seed, cov0, decay_constant, target_acceptance_rate, reg_factor,
                   threshold_sample, nu, eta, adapt_betas) -> SampleResult:
    """The loop of ``sample`` with the per-iteration bookkeeping in ``pv_sampler_propose`` / ``pv_sampler_step``.
    Random numbers are drawn here, in the order the numpy loop draws them."""
    from .. import runtime
    lib = runtime.load_library()
    rs = np.random.RandomState(seed)
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    Q, dim = x0.shape
    C = n_chains
    S = max(C - 1, 1)
    lb = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (dim,)))
    ub = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (dim,)))
    x = np.ascontiguousarray(np.repeat(x0[:, None, :], C, axis=1))
    llh, lpr = (np.array(a, dtype=np.float64) for a in evaluate(x))       # own copies: updated in place
    n_evals = Q * C
    betas = np.ascontiguousarray(np.repeat(near_exponential_decay_betas(C)[None, :], Q, axis=0))
    cov = np.broadcast_to(np.eye(dim) if cov0 is None else np.asarray(cov0, dtype=np.float64), (Q, C, dim, dim)).copy()
    cov = np.ascontiguousarray(_regularize(cov, reg_factor))
    mean_hist, cov_hist, cov_scale = x.copy(), cov.copy(), np.ones((Q, C))
    n_tr = n_samples + 1
    # iteration-major while sampling (contiguous row per iteration), [Q, C, n, ...] in the result
    trace_x, trace_lp, trace_pr = np.empty((n_tr, Q, C, dim)), np.empty((n_tr, Q, C)), np.empty((n_tr, Q, C))
    trace_x[0], trace_lp[0], trace_pr[0] = x, -(llh + lpr), -lpr
    n_acc, n_swap, n_swap_try = np.zeros((Q, C)), np.zeros((Q, S)), np.zeros((Q, S))
    x_new, x_eval, inside = np.empty_like(x), np.empty_like(x), np.empty((Q, C), dtype=np.uint8)
    u_swap = np.zeros((Q, S))
    opts = runtime.SamplerOptions(decay_constant, target_acceptance_rate, reg_factor, nu, eta, int(threshold_sample),
                                  int(bool(adapt_betas)))
    import ctypes
    p_opts = ctypes.addressof(opts)
    ptr = lambda a: a.ctypes.data
    # every array the library updates in place keeps its address for the whole run
    a_propose = (ptr(lb), ptr(ub), ptr(x_new), ptr(inside), ptr(x_eval))
    a_state = (ptr(x), ptr(llh), ptr(lpr), ptr(x_new))
    a_rest = (ptr(u_swap), ptr(betas), ptr(mean_hist), ptr(cov_hist), ptr(cov_scale), ptr(cov), ptr(n_acc), ptr(n_swap),
              ptr(n_swap_try), ptr(trace_x), ptr(trace_lp), ptr(trace_pr))
    p_cov, p_x, p_inside = ptr(cov), ptr(x), ptr(inside)
    for it in range(1, n_samples + 1):
        xi = rs.normal(size=(Q, C, dim))
        if lib.pv_sampler_propose(Q, C, dim, p_cov, p_x, ptr(xi), *a_propose) < 0:
            raise runtime.ParvarB200Error(runtime.last_error(lib))
        llh_new, lpr_new = (np.ascontiguousarray(a, dtype=np.float64) for a in evaluate(x_eval))
        n_evals += Q * C
        u_acc = rs.uniform(size=(Q, C))
        for k in range(C - 1, 0, -1):
            u_swap[:, k - 1] = rs.uniform(size=Q)
        if lib.pv_sampler_step(Q, C, dim, it, n_tr, p_opts, *a_state, ptr(llh_new), ptr(lpr_new), p_inside, ptr(u_acc), *a_rest):
            raise runtime.ParvarB200Error(runtime.last_error(lib))
    return SampleResult(trace_x=np.ascontiguousarray(trace_x.transpose(1, 2, 0, 3)),
                        trace_neglogpost=np.ascontiguousarray(trace_lp.transpose(1, 2, 0)),
                        trace_neglogprior=np.ascontiguousarray(trace_pr.transpose(1, 2, 0)), betas=betas,
                        acceptance_rate=n_acc / n_samples, swap_rate=n_swap / np.maximum(n_swap_try, 1), n_evals=n_evals)


def sample_device(objective, x0: np.ndarray, n_samples: int = 5000, n_chains: int = 4, seed: int = 1234,
                  cov0: Optional[np.ndarray] = None, decay_constant: float = 0.51, target_acceptance_rate: float = 0.234,
                  reg_factor: float = 1e-6, threshold_sample: int = 1, nu: float = 1e3, eta: float = 100.0,
                  adapt_betas: bool = True, rng_block: int = 250, use_graph: bool = True,
                  library_rng: bool = True) -> SampleResult:
    """The loop of ``sample`` resident on the GPU (``pv_sampler_create`` / ``pv_sampler_advance`` / ``pv_sampler_result``):
    chain state, covariances, ladder and trace stay in HBM, one iteration is a recorded CUDA graph (proposal kernel ->
    batched objective -> accept / adapt / swap kernel).  ``objective`` is a ``PetabObjectiveBatch`` (its problem, data,
    priors and bounds are used directly).  Random numbers are drawn here, in the order the numpy loop draws them, and
    uploaded in blocks of ``rng_block`` iterations while the previous block runs - so the device loop walks the same chains as
    ``sample(..., native=False)`` up to round-off.  ``library_rng``: the blocks are drawn by the library's restatement of
    numpy's legacy stream (``pv_sampler_advance_mt``: same numbers, compiled code, interpreter lock released) instead of
    by numpy calls here."""
    import ctypes
    from .. import runtime
    lib = objective.problem._lib or runtime.load_library()     # a run-time compiled model's own library owns its handle
    rs = np.random.RandomState(seed)
    x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)))
    Q, dim = x0.shape
    C, S = int(n_chains), max(int(n_chains) - 1, 1)
    if Q != objective.Q or dim != objective.dim:
        raise ValueError(f"x0 must be [{objective.Q}, {objective.dim}]")
    lb = np.ascontiguousarray(np.broadcast_to(np.asarray(objective.lb, dtype=np.float64), (dim,)))
    ub = np.ascontiguousarray(np.broadcast_to(np.asarray(objective.ub, dtype=np.float64), (dim,)))
    ploc = np.ascontiguousarray(objective.prior_loc, dtype=np.float64)
    pscale = np.ascontiguousarray(np.where(np.isfinite(objective.prior_loc), objective.prior_scale, 1.0), dtype=np.float64)
    opts = runtime.SamplerOptions(decay_constant, target_acceptance_rate, reg_factor, nu, eta, int(threshold_sample),
                                  int(bool(adapt_betas)))
    c0 = None if cov0 is None else np.ascontiguousarray(np.broadcast_to(np.asarray(cov0, dtype=np.float64), (dim, dim)))
    ptr = lambda a: None if a is None else a.ctypes.data   # noqa: E731
    K = int(min(rng_block, n_samples))
    h = lib.pv_sampler_create(objective.problem._h, Q, objective.G, C, int(n_samples), ctypes.addressof(opts), ptr(lb), ptr(ub),
                              ptr(ploc), ptr(pscale), ptr(x0), ptr(c0), K, int(bool(use_graph)))
    if not h:
        raise runtime.ParvarB200Error(f"pv_sampler_create failed: {runtime.last_error(lib)}")
    import time
    tm = {"rng_s": 0.0, "advance_s": 0.0, "result_s": 0.0}
    try:
        done = 0
        mt = runtime.MT19937State.from_random_state(rs) if library_rng else None
        while mt is not None and done < n_samples:
            n = min(K, n_samples - done)
            t_a = time.perf_counter()
            runtime._check(lib.pv_sampler_advance_mt(h, n, ctypes.addressof(mt)), lib)
            tm["advance_s"] += time.perf_counter() - t_a
            done += n
        while done < n_samples:
            n = min(K, n_samples - done)
            t_a = time.perf_counter()
            xi, ua, us = np.empty((n, Q, C, dim)), np.empty((n, Q, C)), np.zeros((n, Q, S))
            for k in range(n):                      # the numpy loop's draw order, iteration by iteration
                xi[k] = rs.normal(size=(Q, C, dim))
                ua[k] = rs.uniform(size=(Q, C))
                for j in range(C - 1, 0, -1):
                    us[k, :, j - 1] = rs.uniform(size=Q)
            t_b = time.perf_counter()
            runtime._check(lib.pv_sampler_advance(h, n, ptr(xi), ptr(ua), ptr(us)), lib)
            tm["rng_s"] += t_b - t_a
            tm["advance_s"] += time.perf_counter() - t_b
            done += n
        t_a = time.perf_counter()
        n_tr = n_samples + 1
        trace_x, trace_lp, trace_pr = np.empty((n_tr, Q, C, dim)), np.empty((n_tr, Q, C)), np.empty((n_tr, Q, C))
        betas, n_acc, n_swap, n_swap_try = np.empty((Q, C)), np.empty((Q, C)), np.empty((Q, S)), np.empty((Q, S))
        n_bad = ctypes.c_int64(0)
        runtime._check(lib.pv_sampler_result(h, ptr(trace_x), ptr(trace_lp), ptr(trace_pr), ptr(betas), ptr(n_acc), ptr(n_swap),
                                             ptr(n_swap_try), ctypes.addressof(n_bad)), lib)
        tm["result_s"] = time.perf_counter() - t_a
    finally:
        lib.pv_sampler_destroy(h)
    objective.n_evals += Q * C * n_tr
    return SampleResult(trace_x=np.ascontiguousarray(trace_x.transpose(1, 2, 0, 3)),
                        trace_neglogpost=np.ascontiguousarray(trace_lp.transpose(1, 2, 0)),
                        trace_neglogprior=np.ascontiguousarray(trace_pr.transpose(1, 2, 0)), betas=betas,
                        acceptance_rate=n_acc / n_samples, swap_rate=n_swap / np.maximum(n_swap_try, 1), n_evals=Q * C * n_tr,
                        timing=tm)


# ---- diagnostics (pyPESTO ``effective_sample_size``: Geweke burn-in + Sokal autocorrelation time) ----------------------
def autocorrelation_time_sokal(chain: np.ndarray) -> float:
    """Integrated autocorrelation time of a 1-D chain with Sokal's adaptive window (as pyPESTO's ``autocorrelation_sokal``)."""
    n = len(chain)
    xc = chain - chain.mean()
    if np.allclose(xc, 0):
        return float(n)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, nfft)
    ac = np.fft.irfft(f * np.conjugate(f))[:n].real
    ac /= ac[0]
    tau = 1.0
    for m in range(1, n):
        tau = 1.0 + 2.0 * np.sum(ac[1:m + 1])
        if m >= 6.0 * tau:           # Sokal window
            break
    return float(max(tau, 1.0))


def geweke_burn_in(chain: np.ndarray, zscore: float = 2.0) -> int:
    """First index (in 5 % steps of the chain) from which first 10 % and last 50 % of the rest agree (Geweke 1992)."""
    n = chain.shape[0]
    for frac in np.linspace(0.0, 0.95, 20):
        s = int(frac * n)
        rest = chain[s:]
        if len(rest) < 20:
            break
        a, b = rest[: max(len(rest) // 10, 2)], rest[len(rest) // 2:]
        z = (a.mean(axis=0) - b.mean(axis=0)) / np.sqrt(a.var(axis=0) / len(a) + b.var(axis=0) / len(b) + 1e-300)
        if np.all(np.abs(z) < zscore):
            return s
    return int(0.5 * n)


def _sokal_tau_batch(chains: np.ndarray) -> np.ndarray:
    """``autocorrelation_time_sokal`` for a batch of chains [n_chains, n] at once (one FFT call, the window search on
    cumulative sums): tau(m) = 1 + 2 sum_{k<=m} rho_k, first m with m >= 6 tau(m)."""
    n = chains.shape[1]
    xc = chains - chains.mean(axis=1, keepdims=True)
    flat = np.all(np.isclose(xc, 0.0), axis=1)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, nfft, axis=1)
    ac = np.fft.irfft(f * np.conjugate(f), axis=1)[:, :n].real
    ac0 = np.where(ac[:, :1] == 0.0, 1.0, ac[:, :1])
    ac = ac / ac0
    tau_m = 1.0 + 2.0 * np.cumsum(ac[:, 1:], axis=1)                 # tau after window m = 1 .. n-1
    m = np.arange(1, n)
    hit = m[None, :] >= 6.0 * tau_m
    first = np.where(hit.any(axis=1), hit.argmax(axis=1), n - 2)
    tau = np.maximum(tau_m[np.arange(len(chains)), first], 1.0)
    return np.where(flat, float(n), tau)


def effective_sample_size(result: SampleResult) -> np.ndarray:
    """ESS of the cold chain of every problem: (n - burn_in) / (1 + max_p tau_p)  (pyPESTO's definition).  Problems that
    share a burn-in length are processed together (``_sokal_tau_batch``)."""
    Q = result.trace_x.shape[0]
    ess = np.zeros(Q)
    burn = np.array([geweke_burn_in(result.trace_x[q, 0]) for q in range(Q)], dtype=int)
    for b in np.unique(burn):
        qs = np.nonzero(burn == b)[0]
        rest = result.trace_x[qs, 0, b:, :]                             # [nq, n - b, dim]
        nq, n, dim = rest.shape
        tau = _sokal_tau_batch(rest.transpose(0, 2, 1).reshape(nq * dim, n)).reshape(nq, dim).max(axis=1)
        ess[qs] = n / (1.0 + tau)
    result.burn_in, result.effective_sample_size = burn, ess
    return ess


def hdi(samples: np.ndarray, prob: float = 0.94) -> tuple[float, float]:
    """Highest-density interval of a 1-D sample (arviz ``az.hdi`` default hdi_prob = 0.94, petab_optimization.py:217-226)."""
    s = np.sort(np.asarray(samples).ravel())
    n = len(s)
    k = int(np.floor(prob * n))
    if k < 1 or k >= n:
        return float(s[0]), float(s[-1])
    widths = s[k:] - s[: n - k]
    i = int(np.argmin(widths))
    return float(s[i]), float(s[i + k])
