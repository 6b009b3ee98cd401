"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding + the single all-reduce."""
import os
import socket

import numpy as np
import pytest

from conftest import PKG, ROOT, pkg


def test_shard_range_partitions_exactly():
    d = pkg("distributed")
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [d.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.shard_range(10, 2, 2)


def _fake_eval_factory(data):
    """logp_i(theta) = -0.5 * (theta0 - data_i)^2 - 0.1 * theta1 ;  analytic gradient."""
    def local_eval(theta, first, n_local):
        C = theta.shape[1] // max(n_local, 1) if n_local else 0
        d = np.tile(data[first:first + n_local], C)
        lp = -0.5 * (theta[0] - d) ** 2 - 0.1 * theta[1]
        g = np.stack([-(theta[0] - d), np.full_like(d, -0.1)])
        return lp, g
    return local_eval


def _worker(rank, world, port, n_ind, n_chains, out):
    import sys
    sys.path.insert(0, str(ROOT))
    import importlib
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module(PKG + ".distributed")
    rs = np.random.RandomState(0)
    data = rs.normal(size=n_ind)
    theta_chain = rs.normal(size=(2, n_chains))                    # shared per chain
    sl = d.ShardedLikelihood(_fake_eval_factory(data), n_chains, n_ind, rank, world)
    theta_local = np.repeat(theta_chain, sl.n_local, axis=1)        # chain-major
    lp, g, shared = sl(theta_local, shared_grad_rows=(0, 1))
    out.put((rank, lp, shared, sl.first, sl.stop))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_likelihood_gloo_world2():
    import torch.multiprocessing as mp
    n_ind, n_chains, world = 11, 3, 2            # ragged: 6 + 5 individuals
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_ind, n_chains, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unsharded truth
    rs = np.random.RandomState(0)
    data = rs.normal(size=n_ind)
    theta_chain = rs.normal(size=(2, n_chains))
    lp_true = np.array([np.sum(-0.5 * (theta_chain[0, c] - data) ** 2 - 0.1 * theta_chain[1, c]) for c in range(n_chains)])
    g0_true = np.array([np.sum(-(theta_chain[0, c] - data)) for c in range(n_chains)])
    spans = sorted((r[3], r[4]) for r in res)
    assert spans == [(0, 6), (6, 11)]
    for rank, lp, shared, first, stop in res:
        np.testing.assert_allclose(lp, lp_true, rtol=1e-12)          # every rank holds the full sums
        np.testing.assert_allclose(shared[:, 0], g0_true, rtol=1e-12)
        np.testing.assert_allclose(shared[:, 1], -0.1 * n_ind, rtol=1e-12)


def test_world1_needs_no_process_group():
    d = pkg("distributed")
    data = np.arange(5.0)
    sl = d.ShardedLikelihood(_fake_eval_factory(data), 2, 5)
    th = np.repeat(np.array([[0.5, 1.5], [1.0, 2.0]]), 5, axis=1)
    lp, g, shared = sl(th, shared_grad_rows=(0,))
    assert lp.shape == (2,) and g.shape == (2, 10) and shared.shape == (2, 1)


def test_assign_longest_first_is_a_balanced_partition():
    """The C5 sweep hands whole time grids to ranks (tools/run_study.py): every unit exactly once, the same answer on
    every rank, the most expensive units on different ranks, loads within one unit of each other."""
    d = pkg("distributed")
    grids = [2, 3, 4, 5, 10, 20, 40, 80, 160]
    costs = {g: 1.0 + g / 160.0 for g in grids}
    for world in (1, 2, 4, 8, 9, 12):
        owner = d.assign_longest_first(costs, world)
        assert owner == d.assign_longest_first(dict(reversed(list(costs.items()))), world)      # input order does not matter
        assert sorted(owner) == grids and all(0 <= r < world for r in owner.values())
        load = [sum(costs[g] for g in grids if owner[g] == r) for r in range(world)]
        used = [x for x in load if x > 0]
        assert len(used) == min(world, len(grids))
        assert max(used) - min(used) <= max(costs.values()) + 1e-12
        top = sorted(grids, key=lambda g: -costs[g])[:min(world, len(grids))]
        assert len({owner[g] for g in top}) == len(top)
    with pytest.raises(ValueError):
        d.assign_longest_first(costs, 0)
