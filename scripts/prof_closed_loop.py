"""closed-loop profile driver: B environments, a few control steps through b2mpc_closed_loop (run under ncu for the launch list)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np, torch
from b200mpc import BatchedMpcSolver, world
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dims = world.MpcDims(20, 5, 2.0)
sc = world.random_scenarios(B, dims, seed=0)
s = BatchedMpcSolver(20, 5, 2.0, max_batch=B)
for k in ("x0", "obst", "goal", "yref"):
    s.set(k, torch.from_numpy(sc[k]).cuda())
s.loop_reset(B); s.reset(B)
s.closed_loop(B, 10, random_move=True, seed=1, tol=-1.0)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); s.closed_loop(B, steps, random_move=True, seed=1, tol=-1.0); e1.record(); torch.cuda.synchronize()
print(f"B={B} {e0.elapsed_time(e1)/steps:.3f} ms per control step, mean qp_iter {s.get_int('qp_iter', B).float().mean().item():.2f}")
