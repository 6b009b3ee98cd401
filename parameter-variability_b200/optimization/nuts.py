"""Batched No-U-Turn sampler for the gradient path (BASELINE.json north_star: "the log-likelihood and its gradient that
the MCMC/NUTS sampler calls at every step").

All chains advance in lock step: in every iteration of the inner loop each chain that is still growing its tree takes ONE
leapfrog step from the end of its trajectory, so one iteration = ONE batched ``logp_and_grad`` evaluation (for the
hierarchical objective: one kernel launch over n_chains x n_individuals trajectories with forward sensitivities).  Chains
whose tree is finished (U-turn, divergence, maximum depth) ride along masked until the slowest chain is done.

Algorithm: multinomial NUTS with the generalised U-turn criterion (Betancourt 2017, as implemented in Stan and NumPyro),
trees built iteratively with O(depth) momentum checkpoints instead of recursion (Phan, Pradhan, Jankowiak 2019, section
"iterative NUTS"), dual-averaging step size (Hoffman & Gelman 2014, Algorithm 5), diagonal mass matrix from the second
half of warm-up.  Restated from the publications: neither PyMC, Stan nor NumPyro is in /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np


@dataclass
class NUTSResult:
    samples: np.ndarray        # [n_chains, n_samples, dim]
    logp: np.ndarray           # [n_chains, n_samples]
    accept_rate: np.ndarray    # [n_chains] mean acceptance statistic after warm-up
    step_size: np.ndarray      # [n_chains]
    inv_mass: np.ndarray       # [n_chains, dim]
    tree_depth: np.ndarray     # [n_chains, n_samples]
    divergences: np.ndarray    # [n_chains]
    n_grad_evals: int          # batched gradient evaluations (= launches of the objective)


def _popcount(n: np.ndarray) -> np.ndarray:
    c = np.zeros_like(n)
    m = n.copy()
    while np.any(m):
        c += m & 1
        m >>= 1
    return c


def _trailing_ones(n: np.ndarray) -> np.ndarray:
    c = np.zeros_like(n)
    m = n.copy()
    go = np.ones_like(n, dtype=bool)
    while np.any(go & (m & 1).astype(bool)):
        hit = go & (m & 1).astype(bool)
        c += hit
        go &= hit
        m >>= 1
    return c


def _is_turning(inv_mass, r_left, r_right, r_sum):
    """Generalised U-turn criterion on the momentum sum of a (sub)tree; symmetric in its two ends."""
    rho = r_sum - 0.5 * (r_left + r_right)
    return (np.sum(inv_mass * r_left * rho, axis=1) <= 0) | (np.sum(inv_mass * r_right * rho, axis=1) <= 0)


def nuts(logp_and_grad: Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]], x0: np.ndarray, n_samples: int = 500,
         n_warmup: int = 500, step_size: float = 0.1, target_accept: float = 0.8, max_depth: int = 8, seed: int = 1234,
         adapt_mass: bool = True, max_energy_error: float = 1000.0) -> NUTSResult:
    """``logp_and_grad(X [C, dim]) -> (logp [C], grad [C, dim])``; ``x0`` [C, dim] (unconstrained coordinates)."""
    rs = np.random.RandomState(seed)
    x = np.array(x0, dtype=np.float64)
    C, dim = x.shape
    lp, g = logp_and_grad(x)
    n_evals = 1
    eps = np.full(C, step_size)
    inv_mass = np.ones((C, dim))
    mu = np.log(10.0 * eps)
    log_eps_bar = np.zeros(C)
    h_bar = np.zeros(C)
    gamma, t0, kappa = 0.05, 10.0, 0.75
    w_n, w_mean, w_m2 = 0, np.zeros((C, dim)), np.zeros((C, dim))
    samples = np.empty((C, n_samples, dim))
    lps = np.empty((C, n_samples))
    depths = np.zeros((C, n_samples), dtype=int)
    acc_sum = np.zeros(C)
    div_count = np.zeros(C, dtype=int)
    ar = np.arange(C)

    for it in range(n_warmup + n_samples):
        warm = it < n_warmup
        r0 = rs.normal(size=(C, dim)) / np.sqrt(inv_mass)
        h0 = -lp + 0.5 * np.sum(inv_mass * r0 * r0, axis=1)
        # whole-tree state
        xl, rl, gl = x.copy(), r0.copy(), g.copy()          # left end
        xr, rr, gr = x.copy(), r0.copy(), g.copy()          # right end
        x_prop, lp_prop, g_prop = x.copy(), lp.copy(), g.copy()
        r_sum = r0.copy()
        log_w = np.zeros(C)                                  # log sum exp(h0 - H) over the tree (root has weight 1)
        depth = np.zeros(C, dtype=int)
        done = np.zeros(C, dtype=bool)
        diverged = np.zeros(C, dtype=bool)
        sum_acc = np.zeros(C)
        n_leaf = np.zeros(C, dtype=int)
        while not np.all(done):
            # ---- start a doubling for every chain that is not done: direction, subtree state ------------------------------
            going_right = rs.uniform(size=C) < 0.5
            n_sub = 1 << depth                               # leaves of the new subtree
            leaf = np.zeros(C, dtype=int)
            xe = np.where(going_right[:, None], xr, xl)      # end from which the subtree grows
            re = np.where(going_right[:, None], rr, rl)
            ge = np.where(going_right[:, None], gr, gl)
            sub_first_r = np.zeros((C, dim))
            sub_r_sum = np.zeros((C, dim))
            sub_log_w = np.full(C, -np.inf)
            sub_x, sub_lp, sub_g = x_prop.copy(), lp_prop.copy(), g_prop.copy()
            sub_turn = np.zeros(C, dtype=bool)
            sub_div = np.zeros(C, dtype=bool)
            r_ck = np.zeros((C, max_depth + 1, dim))
            rs_ck = np.zeros((C, max_depth + 1, dim))
            growing = ~done
            while np.any(growing):
                # ---- one leapfrog step of every growing chain: one batched gradient evaluation ---------------------------
                sgn = np.where(going_right, 1.0, -1.0) * eps
                rh = re + 0.5 * sgn[:, None] * ge
                xn = xe + sgn[:, None] * inv_mass * rh
                lpn, gn = logp_and_grad(np.where(growing[:, None], xn, x))
                n_evals += 1
                bad = ~np.isfinite(lpn) | ~np.all(np.isfinite(gn), axis=1)
                gn = np.where(bad[:, None], 0.0, gn)
                rn = rh + 0.5 * sgn[:, None] * gn
                hn = -lpn + 0.5 * np.sum(inv_mass * rn * rn, axis=1)
                dH = np.where(bad, np.inf, hn - h0)
                dH = np.where(np.isnan(dH), np.inf, dH)
                leaf_div = dH > max_energy_error
                lw = -dH                                       # log weight of the new leaf
                G = growing
                # accept statistic (Stan): min(1, exp(-dH)) averaged over all leaves
                sum_acc += np.where(G, np.minimum(1.0, np.exp(np.minimum(-dH, 0.0))), 0.0)
                n_leaf += G
                # ---- progressive uniform sampling inside the subtree -------------------------------------------------------------
                new_log_w = np.logaddexp(sub_log_w, lw)
                with np.errstate(invalid="ignore"):
                    p_take = np.exp(lw - new_log_w)
                take = G & ~leaf_div & (rs.uniform(size=C) < np.where(np.isnan(p_take), 0.0, p_take))
                sub_x = np.where(take[:, None], xn, sub_x)
                sub_g = np.where(take[:, None], gn, sub_g)
                sub_lp = np.where(take, lpn, sub_lp)
                sub_log_w = np.where(G, new_log_w, sub_log_w)
                first = G & (leaf == 0)
                sub_first_r = np.where(first[:, None], rn, sub_first_r)
                sub_r_sum = np.where(G[:, None], sub_r_sum + rn, sub_r_sum)
                # ---- U-turn inside the subtree via momentum checkpoints (iterative NUTS) -------------------------------------------
                idx_max = _popcount(leaf >> 1)
                n_tr = _trailing_ones(leaf)
                idx_min = idx_max - n_tr + 1
                even = (leaf & 1) == 0
                store = G & even
                r_ck[ar[store], idx_max[store]] = rn[store]
                rs_ck[ar[store], idx_max[store]] = sub_r_sum[store]
                check = G & ~even
                turning = np.zeros(C, dtype=bool)
                for i in range(max_depth, -1, -1):
                    use = check & (i <= idx_max) & (i >= idx_min) & ~turning
                    if not np.any(use):
                        continue
                    sub_sum = sub_r_sum - rs_ck[:, i] + r_ck[:, i]
                    t_i = _is_turning(inv_mass, r_ck[:, i], rn, sub_sum)
                    turning |= use & t_i
                sub_turn |= turning
                sub_div |= G & leaf_div
                # advance the end state
                xe = np.where(G[:, None], xn, xe)
                re = np.where(G[:, None], rn, re)
                ge = np.where(G[:, None], gn, ge)
                leaf = leaf + G
                growing = G & ~sub_turn & ~sub_div & (leaf < n_sub)
            # ---- merge the subtree into the tree ---------------------------------------------------------------------------------
            act = ~done
            ok = act & ~sub_turn & ~sub_div
            # biased progressive sampling between old tree and new subtree
            p_new = np.exp(np.minimum(sub_log_w - log_w, 0.0))
            swap = ok & (rs.uniform(size=C) < p_new)
            x_prop = np.where(swap[:, None], sub_x, x_prop)
            g_prop = np.where(swap[:, None], sub_g, g_prop)
            lp_prop = np.where(swap, sub_lp, lp_prop)
            log_w = np.where(ok, np.logaddexp(log_w, sub_log_w), log_w)
            r_sum = np.where(ok[:, None], r_sum + sub_r_sum, r_sum)
            upd_r = ok & going_right
            upd_l = ok & ~going_right
            xr, rr, gr = (np.where(upd_r[:, None], xe, xr), np.where(upd_r[:, None], re, rr), np.where(upd_r[:, None], ge, gr))
            xl, rl, gl = (np.where(upd_l[:, None], xe, xl), np.where(upd_l[:, None], re, rl), np.where(upd_l[:, None], ge, gl))
            depth = depth + act
            whole_turn = _is_turning(inv_mass, rl, rr, r_sum)
            diverged |= act & sub_div
            done |= act & (sub_turn | sub_div | whole_turn | (depth >= max_depth))
        x, lp, g = x_prop, lp_prop, g_prop
        a = sum_acc / np.maximum(n_leaf, 1)
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
                    inv_mass = (w_n / (w_n + 5.0)) * var + 1e-3 * (5.0 / (w_n + 5.0))
        else:
            k = it - n_warmup
            samples[:, k], lps[:, k], depths[:, k] = x, lp, depth
            acc_sum += a
            div_count += diverged
    return NUTSResult(samples=samples, logp=lps, accept_rate=acc_sum / max(n_samples, 1), step_size=eps, inv_mass=inv_mass,
                      tree_depth=depths, divergences=div_count, n_grad_evals=n_evals)


def nuts_device(H, x0: np.ndarray, n_samples: int = 500, n_warmup: int = 500, step_size: float = 0.1, target_accept: float = 0.8,
                max_depth: int = 8, seed: int = 1234, adapt_mass: bool = True, max_energy_error: float = 1000.0) -> NUTSResult:
    """``nuts`` on the hierarchical posterior with the tree on the GPU (``csrc/pv_nuts.cuh``, entry points ``pv_nuts_begin /
    _start / _drift / _leaf / _merge / _finish``).  ``H`` is a ``HierarchicalObjectiveDevice`` on ONE rank; ``x0`` [C, dim] in the
    order log theta_ci (individual-major), mu_c, log omega_c - the parameterisation of ``tests/test_gpu_estimation.py``.
    Every tree vector, the checkpoints, the adaptation state and the kept samples live in HBM; one leapfrog step of every
    growing chain is drift kernel -> sensitivity kernel + epilogue -> hyper-prior kernel -> leaf kernel.  The host mirrors the
    numpy loop's control flow: it reads C flags after each leaf (is any chain still growing) and each doubling (is every chain
    done) and draws the random numbers in the numpy loop's order - with the same seed both walk the same chains up to round-off."""
    import ctypes
    import torch
    from .. import runtime
    if H.world != 1:
        raise ValueError("nuts_device: the individuals must not be sharded (the U-turn tests are dot products over all coordinates)")
    pb = H.problem
    lib = pb._lib or runtime.load_library()
    C, L, P = H.C, H.n_local, H.P
    dim = L * P + 2 * P
    if C > runtime.NUTS_MAX_CHAINS:
        raise ValueError(f"at most {runtime.NUTS_MAX_CHAINS} chains")
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    if x0.shape != (C, dim):
        raise ValueError(f"x0 must be [{C}, {dim}]")
    dev = H.red.device
    rs = np.random.RandomState(seed)
    V, S, I = ({n: k for k, n in enumerate(names)} for names in (runtime.NUTS_VEC, runtime.NUTS_SC, runtime.NUTS_IC))
    vec = torch.zeros((len(V), C, dim), dtype=torch.float64, device=dev)
    ck = torch.zeros((2, max_depth + 1, C, dim), dtype=torch.float64, device=dev)
    sc = torch.zeros((len(S), C), dtype=torch.float64, device=dev)
    ic = torch.zeros((len(I), C), dtype=torch.int32, device=dev)
    vec[V["X"]] = torch.from_numpy(x0).to(dev)
    vec[V["IM"]] = 1.0
    sc[S["EPS"]] = float(step_size)
    sc[S["DA_MU"]] = float(np.log(10.0 * step_size))
    theta = torch.zeros((P, C, L), dtype=torch.float64, device=dev)
    mu_q = torch.zeros((C, P), dtype=torch.float64, device=dev)
    omega = torch.ones((C, P), dtype=torch.float64, device=dev)
    n_keep = max(int(n_samples), 1)
    samples = torch.zeros((C, n_keep, dim), dtype=torch.float64, device=dev)
    samples_lp = torch.zeros((C, n_keep), dtype=torch.float64, device=dev)
    samples_depth = torch.zeros((C, n_keep), dtype=torch.int32, device=dev)
    flags = torch.zeros((2, C), dtype=torch.int32).pin_memory()      # written by the leaf / merge kernels, read after a stream sync
    flags_np = flags.numpy()
    st = runtime.NUTSState()
    st.n_chains, st.n_par, st.max_depth, st.n_ind, st.n_samples = C, P, int(max_depth), L, n_keep
    st.host_flags = flags.data_ptr()
    for name, t in (("vec", vec), ("ck", ck), ("sc", sc), ("ic", ic), ("theta", theta), ("mu_q", mu_q), ("omega", omega),
                    ("dtheta", H.dtheta), ("red", H.red), ("samples", samples), ("samples_lp", samples_lp),
                    ("samples_depth", samples_depth)):
        setattr(st, name, t.data_ptr())
    sp = ctypes.addressof(st)
    cur = torch.cuda.current_stream()
    stream = cur.cuda_stream
    chk = lambda rc: runtime._check(rc, lib)                                               # noqa: E731
    uni = runtime.NUTSUniforms()
    up = ctypes.addressof(uni)

    def uniforms(values):
        uni.u[:C] = [float(x) for x in values]
        return up

    def leapfrog(u_take):
        chk(lib.pv_nuts_drift(sp, stream))
        H(theta, mu_q, omega)
        chk(lib.pv_nuts_leaf(sp, uniforms(u_take), float(max_energy_error), stream))

    leapfrog(np.zeros(C))                                    # no chain is growing: evaluates the start point
    n_evals = 1
    chk(lib.pv_nuts_finish(sp, 0, 0, int(n_warmup), float(target_accept), int(adapt_mass), 0, stream))
    w_n = 0
    for it in range(int(n_warmup) + int(n_samples)):
        warm = it < n_warmup
        z = torch.from_numpy(rs.normal(size=(C, dim))).to(dev)
        chk(lib.pv_nuts_begin(sp, z.data_ptr(), stream))
        done = np.zeros(C, dtype=bool)
        while not np.all(done):
            chk(lib.pv_nuts_start(sp, uniforms(rs.uniform(size=C)), stream))
            growing = ~done
            while np.any(growing):
                leapfrog(rs.uniform(size=C))
                n_evals += 1
                cur.synchronize()
                growing = flags_np[0] != 0
            chk(lib.pv_nuts_merge(sp, uniforms(rs.uniform(size=C)), stream))
            cur.synchronize()
            done = flags_np[1] != 0
        wn = 0
        if warm and adapt_mass and it >= n_warmup // 2:
            w_n += 1
            wn = w_n
        chk(lib.pv_nuts_finish(sp, 1 if warm else 2, it, int(n_warmup), float(target_accept), int(adapt_mass), wn, stream))
    torch.cuda.current_stream().synchronize()
    return NUTSResult(samples=samples.cpu().numpy()[:, :n_samples], logp=samples_lp.cpu().numpy()[:, :n_samples],
                      accept_rate=sc[S["ACCSUM"]].cpu().numpy() / max(n_samples, 1), step_size=sc[S["EPS"]].cpu().numpy(),
                      inv_mass=vec[V["IM"]].cpu().numpy(), tree_depth=samples_depth.cpu().numpy()[:, :n_samples].astype(int),
                      divergences=ic[I["NDIV"]].cpu().numpy().astype(int), n_grad_evals=n_evals)
