"""Host-side multi-rank logic on CPU: world size 2, gloo backend (what runs over NCCL on the GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faster_b200 import shard


def test_partition_covers_everything():
    for n in (0, 1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            blocks = [shard.partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.partition_sizes(n, world)


def test_select_winners_matches_loop():
    rng = np.random.default_rng(0)
    n_prob, n_dt, n_sig = 9, 6, 11
    cost = rng.uniform(1, 10, (n_prob, n_dt * n_sig))
    cost[rng.uniform(size=cost.shape) < 0.7] = np.inf
    cost[3] = np.inf
    d, s, c = shard.select_winners(cost, n_dt, n_sig)
    for p in range(n_prob):
        grid = cost[p].reshape(n_dt, n_sig)
        exp = next((i for i in range(n_dt) if np.isfinite(grid[i]).any()), -1)
        assert d[p] == exp
        if exp >= 0:
            assert s[p] == int(np.argmin(grid[exp])) and c[p] == grid[exp].min()
        else:
            assert s[p] == -1 and np.isinf(c[p])


def _costs_for(item, cand):
    rng = np.random.default_rng(1000 + item)
    c = rng.uniform(1, 5, cand)
    c[rng.uniform(size=cand) < 0.5] = np.inf
    return c


def _worker(rank, world, port, n_items, cand, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.partition(n_items, world, rank)
    local = torch.from_numpy(np.concatenate([_costs_for(i, cand) for i in range(lo, hi)]) if hi > lo else np.zeros(0))
    full = shard.all_gather_costs(local, n_items, cand)
    q.put((rank, full.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [6, 7])
def test_all_gather_costs_world2(n_items):
    world, cand = 2, 12
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, cand, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expected = np.concatenate([_costs_for(i, cand) for i in range(n_items)])
    for r in range(world):
        assert np.array_equal(got[r], expected)
    d0, s0, c0 = shard.select_winners(got[0].reshape(n_items, cand), 3, 4)
    d1, s1, c1 = shard.select_winners(got[1].reshape(n_items, cand), 3, 4)
    assert np.array_equal(d0, d1) and np.array_equal(s0, s1) and np.array_equal(c0, c1)
