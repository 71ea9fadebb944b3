"""Does HPIPM's `lq_fact` switch change anything the reference's recordings can see?
Replays the reference's 8 recorded settings x 100 seeded episodes (tests/golden/reference_experiments.npz: outcomes written
by src/simulation/experiments.py with the real acados solver) through the CPU ORACLE's closed loop with
lq_fact = 0 (Cholesky of the summed node matrix only), 1 (HPIPM mode BALANCE: switch to the LQ factorisation when the
predictor's linear residual exceeds 1e-5) and 2 (LQ always), and scores every variant PER SEED against the recording
(identical = same control-step count, hit and reached flags).  CPU only (oracle = test infrastructure); one process per core.
usage: python scripts/lq_fact_replay.py [n_seeds] -> one JSON line per variant (kept under profiles/r02_lq_fact_replay.jsonl)"""
import json, os, sys, time
from multiprocessing import Pool
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np
from b200mpc import world
from oracle import pyoracle as po

GX = np.load(os.path.join(ROOT, "tests", "golden", "reference_experiments.npz"))
stamps = sorted(k[5:] for k in GX.files if k.startswith("data_"))


def episode(args):
    """robot_ocp_problem.py:168-277 on the oracle, one seeded episode -> [hit, reached, min_margin, dist, steps, oob]"""
    scen, tf, N, M, qp, seed, lq = args
    cfg = po.default_config(N, M, tf, qp_iter_max=qp, lq_fact=lq)
    dt = tf / N
    obst, noise = world.reference_episode_streams(seed, scen, M, 400, True)
    x0 = np.array([world.X_MIN + 1, world.Y_MIN + 1, np.pi / 4, 0.0, 0.0]); goal = np.array([world.X_MAX - 1, world.Y_MAX - 1])
    x = np.zeros((N + 1, 5)); u = np.zeros((N, 2)); sl = np.zeros((N + 1, M)); su = np.zeros((N + 1, M))
    po.initial_guess(cfg, x0, x, u, sl, su)
    yref = np.zeros((N + 1, 6)); yref[:, :2] = goal
    minm, steps, oob, reached, dist, n_lq = np.inf, 0, 0, 0, 0.0, 0
    for i in range(400):
        p = np.ascontiguousarray(world.predict_trajectory(obst[None], N, dt).reshape(N + 1, 2 * M))
        zl, Zl = po.slack_weights(cfg, x0, goal)
        r = po.rti_solve(cfg, x0, p, yref, zl, Zl, x, u, sl, su)
        n_lq += int(r["stats"][7] > 0)
        u0 = u[0].copy()
        if r["status"] == 4:
            po.initial_guess(cfg, x0, x, u, sl, su)
        x0 = po.gl4_step(x0, u0, dt, sens=False)
        if x0[0] < world.X_MIN or x0[0] > world.X_MAX or x0[1] < world.Y_MIN or x0[1] > world.Y_MAX:
            oob = 1
        obst = world.obstacle_step(obst[None], dt, True, noise[i][None])[0]
        minm = min(minm, np.min(np.sqrt((x0[0] - obst[:, 0]) ** 2 + (x0[1] - obst[:, 1]) ** 2) - 1.2))
        dist = float(np.linalg.norm(x0[:2] - goal))
        if dist <= world.TOL:
            reached = 1
            break
        po.shift(cfg, x, u)
        steps += 1
    return [float(minm <= 0), float(reached), minm, dist, float(steps), float(oob), float(n_lq)]


if __name__ == "__main__":
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    po.build()
    with Pool(os.cpu_count()) as pool:
        for lq in (0, 1, 2):
            t0 = time.time()
            ident, same, agg, lqs = {}, {}, {}, {}
            for st in stamps:
                ref = GX["data_" + st][:n_seeds]
                scen, tf, N, M, qp = GX["spec_" + st]
                jobs = [("RANDOM" if scen == 0 else "EDGE", float(tf), int(N), int(M), int(qp), s, lq) for s in range(n_seeds)]
                ours = np.array(pool.map(episode, jobs, chunksize=2))
                so = (ref[:, 0] == ours[:, 0]) & (ref[:, 1] == ours[:, 1])
                key = f"{'R' if scen == 0 else 'E'}{int(N)}q{int(qp)}"
                ident[key] = float((so & (ref[:, 4] == ours[:, 4])).mean()); same[key] = float(so.mean())
                agg[key] = [round(float(ours[:, 0].mean()), 2), round(float(ours[:, 1].mean()), 2), round(float(ours[:, 4].mean()), 1)]
                lqs[key] = round(float((ours[:, 6] / np.maximum(1, ours[:, 4] + 1)).mean()), 3)
            print(json.dumps({"lq_fact": lq, "identical_mean": round(float(np.mean(list(ident.values()))), 4),
                              "same_outcome_mean": round(float(np.mean(list(same.values()))), 4), "identical": ident,
                              "same_outcome": same, "ours_hit_reached_steps": agg, "share_of_solves_that_switched_to_lq": lqs,
                              "s": round(time.time() - t0, 1)}), flush=True)
