"""debug helper: run both GPU kernels + oracle on the same scenarios; dump to gpurun_out/parity_*.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np, torch
from b200mpc import BatchedMpcSolver, world
from oracle import pyoracle as po

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for (N, M, tf, B, seed, ms) in [(20, 5, 2.0, 2048, 5, True), (10, 3, 1.0, 1024, 5, True), (5, 5, 0.5, 1024, 5, True), (20, 5, 2.0, 2048, 0, False), (30, 8, 3.0, 1024, 5, True)]:
    dims = world.MpcDims(N, M, tf)
    sc = world.random_scenarios(B, dims, seed=seed, moving_start=ms)
    res = {}
    for kern in (1, 2, 3):
        solver = BatchedMpcSolver(N, M, tf, max_batch=B, kernel=kern)
        for name in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su"):
            solver.set(name, torch.from_numpy(sc[name]).cuda())
        solver.solve(B)
        res[f"x{kern}"] = solver.get("x", B).cpu().numpy(); res[f"u{kern}"] = solver.get("u", B).cpu().numpy()
        res[f"it{kern}"] = solver.get_int("qp_iter", B).cpu().numpy(); res[f"stats{kern}"] = solver.get("stats", B).cpu().numpy()
        solver.close()
    cfg = po.default_config(N, M, tf)
    ref = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    st, it, stats = po.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], ref["x"], ref["u"], ref["sl"], ref["su"], nthreads=16)
    np.savez(os.path.join(ROOT, "gpurun_out", f"parity_N{N}_M{M}_s{seed}_{int(ms)}.npz"), ox=ref["x"], ou=ref["u"], oit=it, ostats=stats, **res)
    for kern in (1, 2, 3):
        d = res[f"it{kern}"] - it
        print(N, M, seed, ms, "kernel", kern, "iter match", np.mean(d == 0), "max|d|", np.abs(d).max(), "n |d|>1:", np.sum(np.abs(d) > 1))
