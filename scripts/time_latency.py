"""single-problem latency of the latency kernel (bench.py's leg): p50 / p90 over the first 48 scenarios of the bench workload,
each solved alone, all solves enqueued back to back -- experiment driver (B2MPC_LIB selects a variant build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np, torch
from b200mpc import BatchedMpcSolver, world, capi
if os.environ.get('B2MPC_LIB'):
    capi.LIB_PATH = os.path.join(ROOT, os.environ['B2MPC_LIB'])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dims = world.MpcDims(N, 5, N / 10.0); sc = world.random_scenarios(4096, dims, seed=0)
n = 48
t = {k: torch.from_numpy(np.ascontiguousarray(sc[k][:n])).cuda() for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")}
s = BatchedMpcSolver(N, 5, N / 10.0, max_batch=1, kernel=kern)
pairs, its = [], []
for rep in range(3):
    for i in range(n):
        for k in t: s.set(k, t[k][i:i + 1])
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(1); e1.record()
        if rep > 0: pairs.append((e0, e1))
        if rep == 2: its.append(s.get_int("qp_iter", 1))
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in pairs]); it = np.array([int(x.item()) for x in its])
per = np.median(ms.reshape(2, n), axis=0) / it
print(f"N={N} kernel={s.kernel_name} lib={os.environ.get('B2MPC_LIB','default')}: p50 {np.median(ms):.4f} ms p90 {np.quantile(ms,0.9):.4f} min {ms.min():.4f}; "
      f"median qp_iter {np.median(it)}; us per IPM iteration (median over scenarios) {1e3*np.median(per):.2f}")
