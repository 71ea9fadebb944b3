"""device-timed full-SQP solves of B random scenarios (one launch, per-problem convergence inside the persistent kernel):
prints ms per launch, converged fraction, SQP / IPM iteration statistics -- measurement driver for DESIGN.md"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np
import torch
from b200mpc import BatchedMpcSolver, world
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
lm = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
reps = 4
dims = world.MpcDims(20, 5, 2.0)
sc = world.random_scenarios(B, dims, seed=0)
solver = BatchedMpcSolver(20, 5, 2.0, max_batch=B, lm=lm)
t = {k: torch.from_numpy(sc[k]).cuda() for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")}
for name in ("x0", "p", "yref", "zl", "Zl"):
    solver.set(name, t[name])
ms = []
for _ in range(reps):
    for name in ("x", "u", "sl", "su"):
        solver.set(name, t[name])
    solver.reset_multipliers(B)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); solver.solve_sqp(B, max_iter=max_iter, tol=(tol,) * 4); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = sorted(ms[1:])
st = solver.get_int("status", B).cpu().numpy(); si = solver.get_int("sqp_iter", B).cpu().numpy(); qi = solver.get_int("qp_iter", B).cpu().numpy()
m = ms[len(ms) // 2]
print(json.dumps({"B": B, "max_iter": max_iter, "tol": tol, "lm": lm, "ms_per_launch": m, "nlp_solves_per_s": B / m * 1e3,
                  "qp_solves_per_s": float(si.sum()) / m * 1e3, "converged": float(np.mean(st == 0)), "max_iter_hit": float(np.mean(st == 2)),
                  "qp_fail": float(np.mean(st == 4)), "sqp_iter_mean": float(si.mean()), "sqp_iter_p50": float(np.median(si)),
                  "ipm_iter_per_qp": float(qi.sum() / max(1, si.sum()))}))
