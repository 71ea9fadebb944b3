"""device-timed solves of B random scenarios: prints ms per solve (median of reps) -- experiment driver"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import torch
from b200mpc import BatchedMpcSolver, world, capi
if os.environ.get('B2MPC_LIB'):
    capi.LIB_PATH = os.path.join(ROOT, os.environ['B2MPC_LIB'])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
NN = int(os.environ.get("B2_N", "20"))
dims = world.MpcDims(NN, 5, NN / 10.0)
sc = world.random_scenarios(B, dims, seed=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
solver = BatchedMpcSolver(NN, 5, NN / 10.0, max_batch=B, kernel=kern)
t = {k: torch.from_numpy(sc[k]).cuda() for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")}
for name in ("x0", "p", "yref", "zl", "Zl"):
    solver.set(name, t[name])
ms = []
for _ in range(reps):
    for name in ("x", "u", "sl", "su"):
        solver.set(name, t[name])
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); solver.solve(B); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = sorted(ms[2:])
it = solver.get_int("qp_iter", B).float()
print(f"B={B} kernel={solver.kernel_name} lib={os.environ.get('B2MPC_LIB','default')} split={os.environ.get('B2MPC_EXPERIMENT_SPLIT','-')} ms={ms[len(ms)//2]:.3f} solves/s={B/ms[len(ms)//2]*1e3:.0f} qp_iter_mean={it.mean().item():.3f} sum={int(it.sum().item())}")
