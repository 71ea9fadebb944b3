"""debug build (-DB2_DEBUG_CLOCK): cycles spent in the factorisation / the sequential vector sweeps / the whole solve"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'dynamic-obstacle-avoidance-mpc_b200'))
import torch, numpy as np
from b200mpc import BatchedMpcSolver, world, capi
capi.LIB_PATH = os.path.join(R, os.environ.get('B2MPC_LIB', 'dynamic-obstacle-avoidance-mpc_b200/lib/libb200mpc_dbg.so'))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dims = world.MpcDims(N, 5, N / 10.0); sc = world.random_scenarios(B, dims, seed=0)
for kern in [int(x) for x in os.environ.get("KERNELS", "4").split(",")]:
    s = BatchedMpcSolver(N, 5, N / 10.0, max_batch=B, kernel=kern)
    for k in ("x0","p","yref","zl","Zl","x","u","sl","su"): s.set(k, torch.from_numpy(sc[k]).cuda())
    s.solve(B)
    for k in ("x","u","sl","su"): s.set(k, torch.from_numpy(sc[k]).cuda())
    s.solve(B)
    st = s.get("stats", B).cpu().numpy(); it = s.get_int("qp_iter", B).cpu().numpy().astype(float)
    f, sw, al = st[:, 5] / it, st[:, 6] / it, st[:, 7] / it
    if kern >= 4:
        print("  cta detail per iteration: head %.0f, solve_prep %.0f, riccati_solve (all) %.0f, expand+reduce %.0f, lin_residuals %.0f" % tuple(np.median(st[:, i] / it) for i in range(5)))
    print(f"kernel {kern} B={B} N={N}: per IPM iteration: factorize {np.median(f):.0f} cyc ({np.median(f)/(N+1):.0f}/node), vector sweeps {np.median(sw):.0f} cyc, whole {np.median(al):.0f} cyc; iterations {it[:4]}")
    s.close()
