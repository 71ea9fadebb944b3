"""run the reference's 8 recorded experiment settings (src/simulation/test_data/*_spec.json, without the two
`interpolate_init` pairs) through the batched GPU harness; per-seed outcomes -> gpurun_out/fixture_sweep.npz"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np
from b200mpc import experiments, world
SETTINGS = [("20221031_215846", "RANDOM", 2.0, 20, 100), ("20221031_220136", "EDGE", 2.0, 20, 100),
            ("20221031_220735", "RANDOM", 2.0, 20, 50), ("20221031_220939", "EDGE", 2.0, 20, 50),
            ("20221031_221343", "RANDOM", 2.0, 20, 25), ("20221031_221613", "EDGE", 2.0, 20, 25),
            ("20221031_224515", "RANDOM", 1.0, 10, 50), ("20221031_224642", "EDGE", 1.0, 10, 50)]
out = {}
for stamp, scen, tf, N, qp in SETTINGS:
    t0 = time.time()
    data = experiments.run_scenario(scen, range(100), world.MpcDims(N, 5, tf), qp_iter_max=qp)
    s = experiments.summarize(data)
    print(stamp, scen, tf, N, qp, {k: round(v, 3) for k, v in s.items()}, f"{time.time()-t0:.1f}s", flush=True)
    out[stamp] = data
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "fixture_sweep.npz"), **out)
