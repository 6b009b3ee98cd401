"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), independent trajectories split
contiguously across ranks, and ONE collective - an all-reduce (sum) of the fp64 buffer
``[n_chains, 1 + n_shared]`` = (logp, d logp / d shared parameters) - only when one likelihood spans shards
(SURVEY.md section 8e).  The reference's only parallelism is a process pool over independent PEtab problems
(src/parvar/optimization/petab_optimization.py:381-391); forward simulation (C2) and parameter sweeps (C5)
need no collective at all and simply run on ``shard_range`` slices.

``torch.distributed`` is plumbing here: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``range(n)``: the first ``n % world`` ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def assign_longest_first(costs: dict, world: int) -> dict:
    """Whole work units (e.g. the time grids of the C5 sweep) to ranks: most expensive first, each to the least loaded
    rank (ties: the lowest rank; equally expensive units in key order).  A pure function of ``costs``: every rank computes the same
    assignment without communicating.  Returns {unit: rank}."""
    if world <= 0:
        raise ValueError(f"bad world {world}")
    load = [0.0] * world
    owner = {}
    for unit in sorted(costs, key=lambda k: (-float(costs[k]), k)):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[unit] = r
        load[r] += float(costs[unit])
    return owner


class ShardedLikelihood:
    """log-likelihood of ``n_chains`` parameter sets over ``n_individuals`` individuals sharded across ranks.

    ``local_eval(chain_major_theta[n_cols, n_chains * n_local], first_individual, n_local)`` must return
    ``(logp[n_chains * n_local], grad[n_sens, n_chains * n_local] or None)`` for this rank's individuals - on the
    GPU box that is ``Problem.logp_host`` / ``Problem.logp_grad`` on the rank's device; the CPU tests plug in a
    numpy function.  Per-individual gradient blocks stay local; the chain sums (and the gradient with respect to
    ``shared`` parameters, summed over individuals) are all-reduced in one call.
    """

    def __init__(self, local_eval: Callable, n_chains: int, n_individuals: int, rank: int = 0, world: int = 1,
                 group=None, device: Optional[str] = None):
        self.local_eval = local_eval
        self.n_chains = n_chains
        self.n_individuals = n_individuals
        self.rank, self.world, self.group = rank, world, group
        self.first, self.stop = shard_range(n_individuals, rank, world)
        self.n_local = self.stop - self.first
        self.device = device

    def _all_reduce(self, buf: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return buf
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(buf))
        if self.device is not None:
            t = t.to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def __call__(self, theta_local: np.ndarray, shared_grad_rows: tuple[int, ...] = ()):
        """theta_local [n_cols, n_chains * n_local] (chain-major: column c * n_local + i).
        Returns (logp_chain [n_chains] summed over ALL individuals, grad_local or None,
        shared_grad [n_chains, len(shared_grad_rows)] summed over ALL individuals)."""
        want_grad = True
        logp, grad = self.local_eval(theta_local, self.first, self.n_local)
        if grad is None:
            want_grad = False
        C, L = self.n_chains, self.n_local
        buf = np.zeros((C, 1 + len(shared_grad_rows)))
        if L:
            buf[:, 0] = np.asarray(logp).reshape(C, L).sum(axis=1)
            if want_grad:
                for k, row in enumerate(shared_grad_rows):
                    buf[:, 1 + k] = np.asarray(grad)[row].reshape(C, L).sum(axis=1)
        buf = self._all_reduce(buf)
        return buf[:, 0], (grad if want_grad else None), buf[:, 1:]
