"""Which of the HPIPM / acados options that SURVEY Appendix C could only RECALL do the reference's own recorded acados runs
favour?  Every variant replays the 8 recorded settings (tests/golden/reference_experiments.npz: 100 seeded episodes each)
on the GPU and is scored PER SEED against the recording: identical = same control-step count, hit and reached flags.
usage: option_sweep.py [group]   -> one JSON line per variant"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np
from b200mpc import experiments, world
GX = np.load(os.path.join(ROOT, "tests", "golden", "reference_experiments.npz"))
stamps = sorted(k[5:] for k in GX.files if k.startswith("data_"))
GROUPS = {
    "a": [("baseline", {}), ("mu0=10", dict(mu0=10.0)), ("mu0=100", dict(mu0=100.0)), ("mu0=0.1", dict(mu0=0.1)),
          ("alpha_min=1e-12", dict(alpha_min=1e-12)), ("itref=0", dict(itref_corr_max=0)), ("itref=4", dict(itref_corr_max=4)),
          ("cond_pred_corr=0", dict(cond_pred_corr=0)), ("res_g=1e-8", dict(res_g_max=1e-8)),
          ("res_all=1e-6", dict(res_b_max=1e-6, res_d_max=1e-6, res_m_max=1e-6)), ("thr0=1e-2", dict(thr0=1e-2)), ("thr0=1", dict(thr0=1.0)),
          ("reg_prim=1e-12", dict(reg_prim=1e-12)), ("lm unscaled", dict(lm_scaled_by_dt=0))],
}
GROUPS["b"] = json.loads(os.environ.get("SWEEP_VARIANTS", "[]"))
for name, over in GROUPS[sys.argv[1] if len(sys.argv) > 1 else "a"]:
    t0 = time.time()
    ident = {}; same = {}; dsteps = {}; agg = {}
    for st in stamps:
        ref = GX["data_" + st]
        scen, tf, N, M, qp = GX["spec_" + st]
        ours = experiments.run_scenario("RANDOM" if scen == 0 else "EDGE", range(100), world.MpcDims(int(N), int(M), float(tf)), qp_iter_max=int(qp), **over)
        so = (ref[:, 0] == ours[:, 0]) & (ref[:, 1] == ours[:, 1])
        key = f"{'R' if scen == 0 else 'E'}{int(N)}q{int(qp)}"
        ident[key] = float((so & (ref[:, 4] == ours[:, 4])).mean()); same[key] = float(so.mean())
        dsteps[key] = float(np.median(np.abs(ref[:, 4] - ours[:, 4])))
        agg[key] = [round(float(ours[:, 0].mean()), 2), round(float(ours[:, 1].mean()), 2), round(float(ours[:, 4].mean()), 1)]
    print(json.dumps({"variant": name, "overrides": over, "identical_mean": round(float(np.mean(list(ident.values()))), 4),
                      "same_outcome_mean": round(float(np.mean(list(same.values()))), 4), "identical": ident, "same_outcome": same,
                      "median_abs_dsteps": dsteps, "ours_hit_reached_steps": agg, "s": round(time.time() - t0, 1)}), flush=True)
