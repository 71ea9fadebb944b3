"""GPU vs oracle statistics of the full-SQP mode (FIXED_STEP and MERIT_BACKTRACKING) -- measurement driver for the
thresholds in tests/test_gpu_sqp.py and the table in DESIGN.md"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np
import torch
from b200mpc import BatchedMpcSolver, world
from oracle import pyoracle as po
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dims = world.MpcDims(20, 5, 2.0)
sc = world.random_scenarios(B, dims, seed=14, moving_start=True)
for lm, glob, max_iter, tol in ((2.0, 0, 250, 1e-6), (2.0, 0, 400, 1e-8), (2.0, 1, 250, 1e-6), (0.05, 0, 100, 1e-6), (0.05, 1, 100, 1e-6), (0.0, 1, 100, 1e-6)):
    over = dict(res_g_max=1e-9) if tol < 1e-7 else {}
    s = BatchedMpcSolver(20, 5, 2.0, max_batch=B, lm=lm, nlp_solver=1, **over, nlp_iter_max=max_iter, nlp_tol=(tol,) * 4, nlp_globalization=glob)
    for name in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su"):
        s.set(name, torch.from_numpy(sc[name]).cuda())
    s.reset_multipliers(B)
    s.solve(B)
    out = {k: s.get(k, B).cpu().numpy() for k in ("x", "u", "stats")}
    for k in ("status", "sqp_iter", "qp_iter"):
        out[k] = s.get_int(k, B).cpu().numpy()
    s.close()
    cfg = po.default_config(20, 5, 2.0, lm=lm, **over)
    ref = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    r = po.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], ref["x"], ref["u"], ref["sl"], ref["su"],
                           max_iter=max_iter, tol=(tol,) * 4, nthreads=os.cpu_count(), globalization=glob)
    conv = (r["status"] == 0) & (out["status"] == 0)
    d = np.abs(out["sqp_iter"] - r["sqp_iter"])
    clean = conv & (r["qp_maxed"] == 0)                    # no QP of the run was cut off by qp_iter_max (path dependent from there on)
    same = conv & (d == 0)
    e = np.abs(out["u"] - ref["u"]).reshape(B, -1).max(1) / np.maximum(1.0, np.abs(ref["u"]).reshape(B, -1).max(1))
    print(json.dumps({"lm": lm, "tol": tol, "globalization": glob, "B": B, "gpu_converged": float(np.mean(out["status"] == 0)),
                      "oracle_converged": float(np.mean(r["status"] == 0)), "status_equal": float(np.mean(out["status"] == r["status"])),
                      "sqp_iter_equal_conv": float(np.mean(d[conv] == 0)) if conv.any() else None,
                      "sqp_iter_off_by_gt1_conv": float(np.mean(d[conv] > 1)) if conv.any() else None,
                      "sqp_iter_median_gpu": float(np.median(out["sqp_iter"])), "sqp_iter_median_oracle": float(np.median(r["sqp_iter"])),
                      "err_u_same_iter_max": float(e[same].max()) if same.any() else None, "err_u_conv_max": float(e[conv].max()) if conv.any() else None,
                      "clean_frac": float(np.mean(r["qp_maxed"] == 0)), "clean_conv": int(clean.sum()),
                      "sqp_iter_equal_clean": float(np.mean(d[clean] == 0)) if clean.any() else None,
                      "sqp_iter_off_by_gt1_clean": float(np.mean(d[clean] > 1)) if clean.any() else None,
                      "err_u_clean_max": float(e[clean].max()) if clean.any() else None,
                      "err_u_clean_p99": float(np.percentile(e[clean], 99)) if clean.any() else None,
                      "last_alpha_lt1": float(np.mean(out["stats"][:, 7] < 1.0))}))
