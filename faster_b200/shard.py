"""Multi-GPU sharding of the candidate batch: corridor problems are independent, so they are
partitioned by rank (one process per GPU) and the only exchange is ONE all-gather of per-candidate results
(cost, +inf = infeasible) when the batch outgrows a GPU -- north_star: "a single NCCL all-gather of feasible costs".
Backend-agnostic (NCCL on GPUs, gloo on CPU for the tests); plumbing only, no solver logic lives here.
"""
import numpy as np


def partition(n_items, world, rank):
    """Contiguous, balanced block [lo, hi) of rank `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partition_sizes(n_items, world):
    return [partition(n_items, world, r)[1] - partition(n_items, world, r)[0] for r in range(world)]


def all_gather_costs(local_cost, n_items_total, cand_per_item, group=None):
    """local_cost: torch tensor [n_local_items * cand_per_item] (float64) on this rank's device.
    Returns the full [n_items_total * cand_per_item] tensor in problem order on every rank (one collective).
    Ranks may own different numbers of items; shorter shards are padded to the longest one for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_cost
    sizes = partition_sizes(n_items_total, world)
    longest = max(sizes) * cand_per_item
    send = local_cost
    if send.numel() < longest:
        send = torch.cat([send, torch.full((longest - send.numel(),), float("inf"), dtype=send.dtype, device=send.device)])
    out = torch.empty(world * longest, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(out, send, group=group)
    if all(s == sizes[0] for s in sizes):
        return out
    return torch.cat([out[r * longest: r * longest + sizes[r] * cand_per_item] for r in range(world)])


def select_winners(cost, n_dt, n_sigma):
    """genNewTraj selection (solverGurobi.cpp:445-472) for a stack of problems on the host:
    cost [n_prob, n_dt*n_sigma] (+inf = infeasible, dt-major) -> (dt_index [-1 if none], sigma_index, cost)."""
    c = np.asarray(cost, np.float64).reshape(-1, n_dt, n_sigma)
    feas_dt = np.isfinite(c).any(axis=2)
    dt_idx = np.where(feas_dt.any(axis=1), feas_dt.argmax(axis=1), -1)
    rows = c[np.arange(c.shape[0]), np.maximum(dt_idx, 0)]
    sig_idx = np.where(dt_idx >= 0, rows.argmin(axis=1), -1)
    best = np.where(dt_idx >= 0, rows.min(axis=1), np.inf)
    return dt_idx, sig_idx, best
