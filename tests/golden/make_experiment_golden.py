"""Collect the reference's recorded closed-loop outcomes (src/simulation/test_data/*_experiment_data.csv, written by
src/simulation/experiments.py:12-43 on 2022-10-31 with the real acados solver) into one small fixture:
tests/golden/reference_experiments.npz.  Per setting: data[100, 6] = [hit, reached, min_margin, dist_to_goal, steps,
out_of_bounds] per seed, and the spec (scenario, TF, N_SOLV, N_OBST, QP_ITER).  The two pairs whose spec carries the key
`interpolate_init` (code that is no longer in the repository) are skipped.
usage: python tests/golden/make_experiment_golden.py [/root/reference]"""
import glob, json, os, sys
import numpy as np
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {}
for spec_file in sorted(glob.glob(os.path.join(ref, "src", "simulation", "test_data", "*_experiment_spec.json"))):
    spec = json.load(open(spec_file))
    if "interpolate_init" in spec:
        continue
    stamp = os.path.basename(spec_file)[:15]
    data = np.loadtxt(spec_file.replace("_spec.json", "_data.csv"), delimiter=";")
    assert data.shape == (100, 6) and spec["slack"] and spec["random_move"] and spec["init_guess"] and spec["N_OBST"] == 5
    out[f"data_{stamp}"] = data
    out[f"spec_{stamp}"] = np.array([0.0 if spec["scenario"] == "RANDOM" else 1.0, spec["TF"], spec["N_SOLV"], spec["N_OBST"], spec["QP_ITER"]])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_experiments.npz"), **out)
print(len(out) // 2, "settings")
