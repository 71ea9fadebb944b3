"""throughput of both kernels vs batch size (device-resident inputs, cold-start scenarios)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np, torch
from b200mpc import BatchedMpcSolver, world
dims = world.MpcDims(20, 5, 2.0)
for B in [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]:
    sc = world.random_scenarios(B, dims, seed=0)
    t = {k: torch.from_numpy(sc[k]).cuda() for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")}
    for kern in (2, 1):
        solver = BatchedMpcSolver(20, 5, 2.0, max_batch=B, kernel=kern)
        for name in ("x0", "p", "yref", "zl", "Zl"):
            solver.set(name, t[name])
        ms = []
        for rep in range(3):
            for name in ("x", "u", "sl", "su"):
                solver.set(name, t[name])
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); solver.solve(B); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        it = solver.get_int("qp_iter", B).float().mean().item()
        print(f"B={B:6d} kernel={solver.kernel_name:18s} best {min(ms):9.2f} ms  {B/min(ms)*1e3:10.0f} solves/s  mean_it {it:.1f}  ws {solver.workspace_bytes/2**20:.0f} MiB", flush=True)
        solver.close()
