"""debug helper: run GPU + oracle on the same scenarios and dump both to gpurun_out/parity.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np, torch
from b200mpc import BatchedMpcSolver, world
from oracle import pyoracle as po

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dims = world.MpcDims(20, 5, 2.0)
sc = world.random_scenarios(B, dims, seed=seed, moving_start=True)
solver = BatchedMpcSolver(20, 5, 2.0, max_batch=B)
for name in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su"):
    solver.set(name, torch.from_numpy(sc[name]).cuda())
solver.solve(B)
out = {k: solver.get(k, B).cpu().numpy() for k in ("x", "u", "sl", "su", "stats")}
out["qp_iter"] = solver.get_int("qp_iter", B).cpu().numpy()
cfg = po.default_config(20, 5, 2.0)
ref = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
st, it, stats = po.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], ref["x"], ref["u"], ref["sl"], ref["su"], nthreads=8)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "parity.npz"), gx=out["x"], gu=out["u"], git=out["qp_iter"], gstats=out["stats"],
         ox=ref["x"], ou=ref["u"], oit=it, ostats=stats, ost=st)
eu = np.abs(out["u"] - ref["u"]).reshape(B, -1).max(1)
print("worst problems:", np.argsort(-eu)[:10], eu[np.argsort(-eu)[:10]])
print("their qp_iter gpu/oracle:", out["qp_iter"][np.argsort(-eu)[:10]], it[np.argsort(-eu)[:10]])
