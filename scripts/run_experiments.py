"""python scripts/run_experiments.py [out_dir] [lm_scaled_by_dt]  -- reference protocol on the GPU"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
from b200mpc import experiments, world
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "experiments")
lm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for qp in (50, 100, 25):
    t0 = time.time()
    res = experiments.run_experiment(os.path.join(out, f"lm{lm}_qp{qp}"), qp_iter_max=qp, lm_scaled_by_dt=lm)
    for stamp, data in res:
        print(f"lm_scaled_by_dt={lm} QP_ITER={qp} {stamp.split('_')[-1]}: {experiments.summarize(data)}  ({time.time()-t0:.1f}s)")
