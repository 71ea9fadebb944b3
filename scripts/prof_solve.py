"""minimal driver for ncu: a few solves of B random scenarios (no oracle, no host copies)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import torch
from b200mpc import BatchedMpcSolver, world
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dims = world.MpcDims(20, 5, 2.0)
sc = world.random_scenarios(B, dims, seed=0)
solver = BatchedMpcSolver(20, 5, 2.0, max_batch=B, kernel=kern)
t = {k: torch.from_numpy(sc[k]).cuda() for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")}
for name in ("x0", "p", "yref", "zl", "Zl"):
    solver.set(name, t[name])
for _ in range(reps):
    for name in ("x", "u", "sl", "su"):
        solver.set(name, t[name])
    solver.solve(B)
torch.cuda.synchronize()
print("qp_iter mean", solver.get_int("qp_iter", B).float().mean().item())
