"""Environment sharding across GPUs.  The batch of MPC problems is embarrassingly parallel:
scenario b of a global batch of size B_total goes to rank b // (B_total / world) (contiguous chunks), every
rank solves its chunk with its own handle on its own GPU, and the only communication is one reduction of a
few scalars at the end (solve count, iteration sum, failures: SUM; elapsed time: MAX).  No data-path collective.
Works on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int):
    """[lo, hi) of the contiguous chunk of `rank`; chunks differ by at most one element"""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_dict(arrays: dict, rank: int, world: int):
    """slice every array of a scenario dict along the batch dimension"""
    total = next(iter(arrays.values())).shape[0]
    lo, hi = shard_bounds(total, rank, world)
    return {k: np.ascontiguousarray(v[lo:hi]) for k, v in arrays.items()}


def reduce_stats(sums: dict, maxs: dict, device=None):
    """all-reduce {name: number}: `sums` with SUM, `maxs` with MAX; identity when not distributed"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(sums), dict(maxs)
    ks, km = sorted(sums), sorted(maxs)
    ts = torch.tensor([float(sums[k]) for k in ks], dtype=torch.float64, device=device)
    tm = torch.tensor([float(maxs[k]) for k in km], dtype=torch.float64, device=device)
    dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    return dict(zip(ks, ts.tolist())), dict(zip(km, tm.tolist()))


def gather_rows(local: np.ndarray, device=None):
    """all-gather per-problem rows (possibly different counts per rank) -> global array on every rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [None] * world
    dist.all_gather_object(counts, int(local.shape[0]))
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(device) if device is not None else torch.from_numpy(pad)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)], axis=0)
