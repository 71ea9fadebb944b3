"""Experiment harness with the reference's protocol and file formats (src/simulation/experiments.py:12-43):
scenarios x 100 seeds x <= 400 control steps -> `<timestamp>_experiment_data.csv` (';'-separated, 6 columns
[hit, reached, min_margin, dist_to_goal, steps, out_of_bounds]) + `<timestamp>_experiment_spec.json`.
All seeds of a scenario run as ONE batch on the GPU; obstacle scenarios and noise streams are the ones
`np.random.seed(i)` gives the reference (world.reference_episode_streams)."""
from __future__ import annotations

import json
import os
from datetime import datetime

import numpy as np

from . import world
from .closed_loop import BatchedRobotOcp


def run_scenario(scenario, seeds, dims=world.MpcDims(), max_steps=400, random_move=True, init_guess=True,
                 qp_iter_max=world.QP_ITER, start=(world.X_MIN + 1, world.Y_MIN + 1, np.pi / 4, 0.0, 0.0),
                 goal=(world.X_MAX - 1, world.Y_MAX - 1), device=0, **solver_overrides):
    """-> data[len(seeds), 6] in the reference's column order (robot_ocp_problem.py:277, experiments.py:36)."""
    streams = [world.reference_episode_streams(s, scenario, dims.M, max_steps, random_move) for s in seeds]
    obst = np.stack([o for o, _ in streams])
    noise = np.ascontiguousarray(np.stack([n for _, n in streams], axis=1))       # [steps, B, M, 2]
    ocp = BatchedRobotOcp(np.array(start, float), np.array(goal, float), obst, dims, slack=True,
                          init_guess_when_error=init_guess, random_move=random_move, noise=noise, device=device,
                          qp_iter_max=qp_iter_max, **solver_overrides)
    x, hit, reached, min_margin, dist, steps, oob = ocp.step(max_steps)
    ocp.solver.close()
    return np.stack([hit.astype(float), reached.astype(float), min_margin, dist, steps.astype(float), oob.astype(float)], axis=1)


def run_experiment(out_dir="test_data/new", scenarios=("RANDOM", "EDGE"), n_seeds=100, dims=world.MpcDims(),
                   qp_iter_max=world.QP_ITER, **kw):
    os.makedirs(out_dir, exist_ok=True)
    written = []
    for s in scenarios:
        data = run_scenario(s, range(n_seeds), dims, qp_iter_max=qp_iter_max, **kw)
        t = datetime.now()
        stamp = f"{t.year:04d}{t.month:02d}{t.day:02d}_{t.hour:02d}{t.minute:02d}{t.second:02d}_{s}"
        spec = {"slack": True, "random_move": True, "init_guess": True, "scenario": s, "TF": dims.tf,
                "N_SOLV": dims.N, "N_OBST": dims.M, "QP_ITER": qp_iter_max}
        np.savetxt(os.path.join(out_dir, f"{stamp}_experiment_data.csv"), data, delimiter=";")
        with open(os.path.join(out_dir, f"{stamp}_experiment_spec.json"), "w") as fh:
            json.dump(spec, fh)
        written.append((stamp, data))
    return written


def summarize(data):
    """evaluate_experiments.py:21-33 ratios"""
    return dict(hit=float(data[:, 0].mean()), reached=float(data[:, 1].mean()), mean_steps=float(data[:, 4].mean()),
                oob=float(data[:, 5].mean()), n=int(data.shape[0]))
