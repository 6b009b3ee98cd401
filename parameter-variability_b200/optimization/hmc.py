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


def hmc_device(H, x0: np.ndarray, n_samples: int = 500, n_warmup: int = 500, n_leapfrog: int = 16, step_size: float = 0.05,
               target_accept: float = 0.8, seed: int = 1234, adapt_mass: bool = True, max_energy_error: float = 1000.0,
               keep_theta: bool = True) -> HMCResult:
    """``hmc`` on the hierarchical posterior with the sampler's state on the GPU (``csrc/pv_hmc.cuh``, entry points
    ``pv_hmc_begin / _drift / _kick / _accept``): ``H`` is a ``HierarchicalObjectiveDevice`` (its buffers ``red`` / ``dtheta``
    are the objective outputs the kick kernel reads), every leapfrog step is drift kernel -> sensitivity kernel + epilogue
    [-> all-reduce] -> hyper-prior kernel -> kick kernel on torch's current stream, nothing is copied or synchronised until
    the kept samples are read at the end.

    The unknowns of chain c, in the order of ``x0`` [C, n_individuals * P + 2 P] (all individuals, also when they are sharded
    over ranks): log theta_ci for individual i = 0.., parameter p = 0..P-1 (individual-major), then mu_c [P], then
    log omega_c [P] - the log-scale parameterisation of ``tests/test_gpu_estimation.py``.  Random numbers are drawn here, in
    the order ``hmc`` draws them, and uploaded in blocks of 128 iterations; with the same seed both walk the same chains up to
    round-off.
    With sharded individuals (``H.world > 1``) every rank passes the same ``x0`` and seed; the returned theta block holds
    this rank's individuals (NaN elsewhere), the hyper-parameters are complete on every rank."""
    import ctypes
    import torch
    from .. import runtime
    pb = H.problem
    lib = pb._lib or runtime.load_library()
    C, L, P, n_tot, first = H.C, H.n_local, H.P, H.n_ind, H.first
    dim = n_tot * P + 2 * P
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    if x0.shape != (C, dim):
        raise ValueError(f"x0 must be [{C}, {dim}]")
    dev = H.red.device
    n_iter = int(n_warmup) + int(n_samples)
    rs = np.random.RandomState(seed)
    # Random numbers in blocks of 128 iterations (bounded memory: a block of 8 chains x 2004 unknowns is 16 MB).  The next block is
    # drawn while the GPU works through the kernels queued for the current one; its upload is ordered on the same stream behind
    # them, so it never overwrites numbers a queued kernel still reads.
    block = max(1, min(n_iter, 128))
    z_d = torch.empty((block, C, dim), dtype=torch.float64, device=dev)
    u_d = torch.empty((block, C), dtype=torch.float64, device=dev)
    Ls: list[int] = []

    def draw_block(first_it):                # the draw order of hmc(): momenta, trajectory length, acceptance uniforms
        n = min(block, n_iter - first_it)
        z, u = np.empty((n, C, dim)), np.empty((n, C))
        for j in range(n):
            z[j] = rs.normal(size=(C, dim))
            Ls.append(max(1, int(round(n_leapfrog * rs.uniform(0.8, 1.2)))))
            u[j] = rs.uniform(size=C)
        z_d[:n].copy_(torch.from_numpy(z))
        u_d[:n].copy_(torch.from_numpy(u))

    def t64(*shape, fill=0.0):
        return torch.full(shape, fill, dtype=torch.float64, device=dev)

    th_shape, hy_shape = (P, C, L), (C, 2 * P)
    buf = {"lth": t64(*th_shape), "mu": t64(C, P), "lom": t64(C, P), "g_lth": t64(*th_shape), "g_hyp": t64(*hy_shape), "lp": t64(C),
           "lth_q": t64(*th_shape), "mu_q": t64(C, P), "lom_q": t64(C, P), "g_lth_q": t64(*th_shape), "g_hyp_q": t64(*hy_shape),
           "p_lth": t64(*th_shape), "p_hyp": t64(*hy_shape), "im_lth": t64(*th_shape, fill=1.0), "im_hyp": t64(*hy_shape, fill=1.0),
           "eps": t64(C, fill=float(step_size)), "theta": t64(*th_shape), "omega": t64(C, P), "dtheta": H.dtheta, "red": H.red,
           "energy": t64(C, 4), "ok": torch.ones(C, dtype=torch.int32, device=dev), "da": t64(C, 4),
           "w_mean_lth": t64(*th_shape), "w_m2_lth": t64(*th_shape), "w_mean_hyp": t64(*hy_shape), "w_m2_hyp": t64(*hy_shape),
           "samples_lth": t64(max(n_samples, 1), *th_shape) if keep_theta else None,
           "samples_hyp": t64(max(n_samples, 1), C, 2 * P), "samples_lp": t64(max(n_samples, 1), C), "acc_sum": t64(C),
           "n_div": torch.zeros(C, dtype=torch.int32, device=dev)}
    lth0 = x0[:, :n_tot * P].reshape(C, n_tot, P)[:, first:first + L]                       # [C, L, P]
    for name in ("lth", "lth_q"):
        buf[name].copy_(torch.from_numpy(np.ascontiguousarray(lth0.transpose(2, 0, 1))))
    for name, col in (("mu", 0), ("mu_q", 0), ("lom", P), ("lom_q", P)):
        buf[name].copy_(torch.from_numpy(np.ascontiguousarray(x0[:, n_tot * P + col:n_tot * P + col + P])))
    buf["da"][:, 0] = float(np.log(10.0 * step_size))
    st = runtime.HMCState()
    st.n_chains, st.n_par, st.own_hyper, st.n_local, st.n_total, st.first = C, P, int(first == 0), L, n_tot, first
    for name in runtime.HMCState.POINTERS:
        setattr(st, name, None if buf[name] is None else buf[name].data_ptr())
    sp = ctypes.addressof(st)
    stream = torch.cuda.current_stream().cuda_stream
    chk = lambda rc: runtime._check(rc, lib)                                               # noqa: E731
    n_evals = 0

    def leapfrog(scale, last):
        chk(lib.pv_hmc_drift(sp, float(scale), stream))
        H(buf["theta"], buf["mu_q"], buf["omega"])          # -> H.dtheta, H.red (all-reduced and with the hyper-priors)
        chk(lib.pv_hmc_kick(sp, float(scale), int(last), stream))

    def reduce_energy():
        if H.world > 1:
            import torch.distributed as dist
            dist.all_reduce(buf["energy"], op=dist.ReduceOp.SUM, group=H.group)

    leapfrog(0.0, True)                                       # log density and gradient of the start point
    n_evals += 1
    reduce_energy()
    chk(lib.pv_hmc_accept(sp, 0, 0, int(n_warmup), None, float(target_accept), float(max_energy_error), int(adapt_mass), 0, stream))
    w_n = 0
    if n_iter:
        draw_block(0)
    for it in range(n_iter):
        warm = it < n_warmup
        j = it % block
        chk(lib.pv_hmc_begin(sp, z_d[j].data_ptr(), stream))
        for k in range(Ls[it]):
            leapfrog(1.0, k == Ls[it] - 1)
        n_evals += Ls[it]
        reduce_energy()
        wn = 0
        if warm and adapt_mass and it >= n_warmup // 2:
            w_n += 1
            wn = w_n
        chk(lib.pv_hmc_accept(sp, 1 if warm else 2, it, int(n_warmup), u_d[j].data_ptr(), float(target_accept),
                              float(max_energy_error), int(adapt_mass), wn, stream))
        if j == block - 1 and it + 1 < n_iter:
            draw_block(it + 1)
    torch.cuda.current_stream().synchronize()
    samples = np.full((C, n_samples, dim), np.nan)
    if n_samples:
        hyp = buf["samples_hyp"].cpu().numpy()                                             # [n, C, 2P]
        samples[:, :, n_tot * P:] = hyp.transpose(1, 0, 2)
        if keep_theta:
            sl = buf["samples_lth"].cpu().numpy()                                          # [n, P, C, L]
            samples[:, :, first * P:(first + L) * P] = sl.transpose(2, 0, 3, 1).reshape(C, n_samples, L * P)
    inv_mass = np.full((C, dim), np.nan)
    inv_mass[:, first * P:(first + L) * P] = buf["im_lth"].cpu().numpy().transpose(1, 2, 0).reshape(C, L * P)
    inv_mass[:, n_tot * P:] = buf["im_hyp"].cpu().numpy()
    return HMCResult(samples=samples, logp=buf["samples_lp"].cpu().numpy().T[:, :n_samples].copy(),
                     accept_rate=buf["acc_sum"].cpu().numpy() / max(n_samples, 1), step_size=buf["eps"].cpu().numpy(),
                     inv_mass=inv_mass, n_grad_evals=n_evals, divergences=buf["n_div"].cpu().numpy().astype(int))
