"""Multi-GPU path on CPU: world_size-2 gloo processes shard a seeded scenario batch, solve their chunks (with
the oracle standing in for the device -- the sharding logic is device-agnostic), reduce the statistics and
gather the controls; results are bit-identical to the unsharded run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200mpc import sharding, world


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 4096, 262144, 1001):
        for w in (1, 2, 3, 4, 8):
            b = [sharding.shard_bounds(total, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _solve(sc):
    from oracle import pyoracle as po
    cfg = po.default_config(10, 5, 1.0)
    x, u, sl, su = (sc[k].copy() for k in ("x", "u", "sl", "su"))
    st, it, _ = po.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], x, u, sl, su, nthreads=1)
    return u[:, 0].copy(), st, it


def _worker(rank, wsize, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=wsize)
    sc = world.random_scenarios(B, world.MpcDims(10, 5, 1.0), seed=7)
    keys = ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")
    mine = sharding.shard_dict({k: sc[k] for k in keys}, rank, wsize)
    u0, st, it = _solve(mine)
    sums, maxs = sharding.reduce_stats({"solves": len(st), "qp_iter": int(it.sum()), "fail": int((st != 0).sum())}, {"time": 1.0 + rank})
    allu = sharding.gather_rows(u0)
    if rank == 0:
        q.put((sums, maxs, allu))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_equals_single_process():
    B = 37                                   # odd on purpose: ragged shards
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums, maxs, allu = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = world.random_scenarios(B, world.MpcDims(10, 5, 1.0), seed=7)
    u_ref, st, it = _solve(sc)
    assert np.array_equal(allu, u_ref)                        # bit-identical whatever the shard count
    assert sums == {"fail": float((st != 0).sum()), "qp_iter": float(it.sum()), "solves": float(B)}
    assert maxs == {"time": 2.0}
