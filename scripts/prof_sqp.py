"""minimal driver for compute-sanitizer / ncu on the full-SQP instantiation: one solve_sqp launch of B random scenarios
usage: prof_sqp.py <B> <kernel 0|3> <max_iter> <globalization 0|1> [lm]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import torch
from b200mpc import BatchedMpcSolver, world
B = int(sys.argv[1]); kern = int(sys.argv[2]); max_iter = int(sys.argv[3]); glob = int(sys.argv[4])
lm = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
dims = world.MpcDims(20, 5, 2.0)
sc = world.random_scenarios(B, dims, seed=0)
s = BatchedMpcSolver(20, 5, 2.0, max_batch=B, kernel=kern, lm=lm, nlp_solver=1, nlp_iter_max=max_iter, nlp_globalization=glob)
for name in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su"):
    s.set(name, torch.from_numpy(sc[name]).cuda())
s.reset_multipliers(B)
s.solve(B)
torch.cuda.synchronize()
print("status", torch.bincount(s.get_int("status", B), minlength=5).tolist(), "sqp_iter mean", s.get_int("sqp_iter", B).float().mean().item())
