"""GPU tests of the on-device closed loop around the solve against the golden vectors
recorded from the reference's own Python code and against the oracle-driven loop; plus the reference's
experiment protocol / file format and its closed-loop statistics band."""
import json
import os

import numpy as np
import pytest

from b200mpc import world

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
DT = 0.1


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


@pytest.mark.parametrize("scen", ["RANDOM", "EDGE"])
def test_obstacle_kernels_match_reference_golden(torch_cuda, scen):
    torch = torch_cuda
    from b200mpc import BatchedMpcSolver
    init, traj, pred0, predm = (G[f"{k}_{scen}"] for k in ("obst_init", "obst_traj", "obst_pred0", "obst_predmid"))
    B, steps = init.shape[0], traj.shape[1]
    s = BatchedMpcSolver(20, 5, 2.0, max_batch=B)
    s.loop_reset(B)
    s.set("obst", init); s.set("goal", np.tile([100.0, 100.0], (B, 1)))      # far goal: never "done"
    s.set("x0", np.zeros((B, 5)))
    s.obst_predict(B, quirk=True)
    p = s.get("p", B, host=True).reshape(B, 21, 5, 2)
    assert np.allclose(np.transpose(p, (0, 2, 1, 3)), pred0, rtol=0, atol=1e-13)
    noise = np.stack([world.reference_episode_streams(seed, scen, 5, steps, True)[1] for seed in range(B)], axis=1)   # [steps,B,M,2]
    for k in range(steps):
        s.loop_post(B, torch.from_numpy(np.ascontiguousarray(noise[k])).cuda(), random_move=True, init_guess_when_error=False, tol=0.15)
        o = s.get("obst", B, host=True)
        assert np.allclose(o, traj[:, k], rtol=0, atol=1e-13), k
        if k == steps // 2:
            s.obst_predict(B, quirk=True)
            p = s.get("p", B, host=True).reshape(B, 21, 5, 2)
            assert np.allclose(np.transpose(p, (0, 2, 1, 3)), predm, rtol=0, atol=1e-13)
    assert (s.get_int("steps", B, host=True) == steps).all()
    s.close()


def test_slack_reset_shift_kernels_match_reference_golden(torch_cuda):
    from b200mpc import BatchedMpcSolver
    N = int(G["ocp_N"][0])
    s = BatchedMpcSolver(N, 5, 0.1 * N, max_batch=4)
    s.set("x0", np.tile(G["slack_x0"], (4, 1))); s.set("goal", np.tile(G["slack_goal"], (4, 1)))
    s.prepare_slack(4)
    assert np.allclose(s.get("zl", 4, host=True)[2], G["slack_zl"], rtol=1e-15, atol=0)
    assert np.allclose(s.get("Zl", 4, host=True)[1], G["slack_Zl"], rtol=1e-15, atol=0)
    s.set("x0", np.tile([1.0, -2.0, 0.3, 4.0, 5.0], (4, 1)))
    s.reset(4)
    assert np.array_equal(s.get("x0", 4, host=True)[3], G["guess_x0_after"])
    assert np.array_equal(s.get("x", 4, host=True)[0], G["guess_x"]) and np.array_equal(s.get("u", 4, host=True)[0], G["guess_u"])
    s.set("x", np.tile(G["shift_X_in"], (4, 1, 1))); s.set("u", np.tile(G["shift_U_in"], (4, 1, 1)))
    s.shift(4)
    assert np.array_equal(s.get("x", 4, host=True)[1], G["shift_X_out"]) and np.array_equal(s.get("u", 4, host=True)[1], G["shift_U_out"])
    s.close()


def _oracle_episode(oracle, cfg, obst, noise, x0, goal, steps, init_guess_when_error=True):
    """the reference's closed loop (robot_ocp_problem.py:168-277) restated on the oracle, one environment"""
    N, M = cfg.N, cfg.M
    dt = cfg.tf / N
    x0 = x0.copy(); obst = obst.copy()
    x = np.zeros((N + 1, 5)); u = np.zeros((N, 2)); sl = np.zeros((N + 1, M)); su = np.zeros((N + 1, M))
    oracle.initial_guess(cfg, x0, x, u, sl, su)
    yref = np.zeros((N + 1, 6)); yref[:, :2] = goal
    log = []
    minm = np.inf
    for i in range(steps):
        p = np.ascontiguousarray(world.predict_trajectory(obst, N, dt).reshape(N + 1, 2 * M))
        zl, Zl = oracle.slack_weights(cfg, x0, goal)
        r = oracle.rti_solve(cfg, x0, p, yref, zl, Zl, x, u, sl, su)
        u0 = u[0].copy()
        if r["status"] == 4 and init_guess_when_error:
            oracle.initial_guess(cfg, x0, x, u, sl, su)
        x0 = oracle.gl4_step(x0, u0, dt, sens=False)
        obst = world.obstacle_step(obst, dt, True, noise[i])
        minm = min(minm, np.min(np.sqrt((x0[0] - obst[:, 0]) ** 2 + (x0[1] - obst[:, 1]) ** 2) - 1.2))
        log.append((x0.copy(), u0, r["qp_iter"], minm))
        if np.linalg.norm(x0[:2] - goal) <= world.TOL:
            break
        oracle.shift(cfg, x, u)
    return log


def test_batched_closed_loop_tracks_oracle_loop(torch_cuda, oracle):
    """free-running (no re-synchronisation) closed loop of 6 reference-seeded episodes, 20 control steps"""
    from b200mpc.closed_loop import BatchedRobotOcp
    dims = world.MpcDims(20, 5, 2.0)
    seeds, steps = list(range(6)), 20
    streams = [world.reference_episode_streams(s, "RANDOM", 5, steps, True) for s in seeds]
    obst = np.stack([o for o, _ in streams]); noise = np.ascontiguousarray(np.stack([n for _, n in streams], axis=1))
    start = np.array([-6.0, -6.0, np.pi / 4, 0.0, 0.0]); goal = np.array([6.0, 6.0])
    ocp = BatchedRobotOcp(start, goal, obst, dims, init_guess_when_error=True, random_move=True, noise=noise)
    cfg = oracle.default_config(20, 5, 2.0)
    logs = [_oracle_episode(oracle, cfg, obst[b], noise[:, b], start, goal, steps) for b in range(len(seeds))]
    s = ocp.solver
    ocp.set_initial_guess()
    tracking = np.ones(len(seeds), bool)      # an environment is compared until its first non-converged QP:
    checked = 0                               # from there on the two (chaotic) closed loops legitimately separate
    for i in range(steps):
        s.obst_predict(len(seeds)); s.prepare_slack(len(seeds)); s.solve(len(seeds))
        s.loop_post(len(seeds), ocp._noise[i], True, True, world.TOL)
        x0 = s.get("x0", len(seeds), host=True); u0 = s.get("u0", len(seeds), host=True)
        it = s.get_int("qp_iter", len(seeds), host=True); mm = s.get("margin", len(seeds), host=True)[:, 0]
        for b in range(len(seeds)):
            if i >= len(logs[b]):
                tracking[b] = False
                continue
            xo, uo, ito, mo = logs[b][i]
            if ito >= 50 or it[b] >= 50:
                tracking[b] = False
            if not tracking[b]:
                continue
            checked += 1
            assert np.abs(u0[b] - uo).max() < 1e-5 * max(1.0, np.abs(uo).max()), (i, b)
            assert np.abs(x0[b] - xo).max() < 1e-5, (i, b)
            assert abs(int(it[b]) - ito) <= 1 and abs(mm[b] - mo) < 1e-5
    assert checked >= 40
    s.close()


def test_experiment_protocol_file_format_and_statistics_band(torch_cuda, tmp_path):
    """src/simulation/experiments.py protocol (2 scenarios x 100 seeds x <= 400 steps, Tf=2, N=20, QP_ITER=50);
    the reference's own result files for this setting (test_data/20221031_220735/_220939) give
    RANDOM hit 18 % / reached 94 % / 120.8 steps and EDGE 16 % / 84 % / 180.2 steps; closed loops are chaotic in
    the seed and the acados version is unpinned, so only the band is asserted."""
    from b200mpc import experiments
    res = experiments.run_experiment(str(tmp_path), n_seeds=100)
    assert len(res) == 2
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 4 and all(f.endswith(("_experiment_data.csv", "_experiment_spec.json")) for f in files)
    ref = {"RANDOM": (0.18, 0.94, 120.8), "EDGE": (0.16, 0.84, 180.2)}
    for stamp, data in res:
        scen = stamp.split("_")[-1]
        disk = np.loadtxt(os.path.join(tmp_path, f"{stamp}_experiment_data.csv"), delimiter=";")
        assert disk.shape == (100, 6) and np.allclose(disk, data)
        spec = json.load(open(os.path.join(tmp_path, f"{stamp}_experiment_spec.json")))
        assert spec == {"slack": True, "random_move": True, "init_guess": True, "scenario": scen, "TF": 2.0, "N_SOLV": 20,
                        "N_OBST": 5, "QP_ITER": 50}
        s = experiments.summarize(data)
        hit, reached, steps = ref[scen]
        assert 0.03 <= s["hit"] <= 0.35, s
        assert reached - 0.12 <= s["reached"] <= 1.0, s
        assert 0.8 * steps <= s["mean_steps"] <= 1.2 * steps, s
        assert set(np.unique(data[:, 0])) <= {0.0, 1.0} and (data[:, 4] <= 400).all()


def test_rl_env_api_shape_and_progress(torch_cuda):
    """src/train.py:10-29 API: reset() -> state[B, 3(M+1)], step(action[B,2]) -> (state, reward, done, info);
    a policy that proposes the final goal as sub-goal drives the robots towards it through the GPU MPC."""
    from b200mpc.rl_env import BatchedMpcEnv
    env = BatchedMpcEnv(64, mpc_steps=5, max_env_steps=30, seed=3)
    state = env.reset()
    assert state.shape == (64, 18) and env.state_dim == 18 and env.action_dim == 2
    d0 = np.linalg.norm(state[:, :2] - env.goal, axis=1)
    total = np.zeros(64)
    for _ in range(12):
        state, reward, done, info = env.step(env.goal)
        assert state.shape == (64, 18) and reward.shape == (64,) and done.shape == (64,) and done.dtype == bool
        total += reward
    d1 = np.linalg.norm(state[:, :2] - env.goal, axis=1)
    assert np.median(d1) < 0.6 * np.median(d0)              # 60 control steps of 0.1 s: clearly closer to the goal
    assert np.isfinite(total).all()
    # same seed -> same rollout (deterministic apart from the torch noise stream, disabled here)
    env2 = BatchedMpcEnv(8, mpc_steps=3, seed=5, random_move=False); s2 = env2.reset(); a = env2.step(env2.goal)[0]
    env3 = BatchedMpcEnv(8, mpc_steps=3, seed=5, random_move=False); s3 = env3.reset(); b = env3.step(env3.goal)[0]
    assert np.array_equal(s2, s3) and np.array_equal(a, b)
    env.close(); env2.close(); env3.close()


GX = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_experiments.npz"))


@pytest.mark.parametrize("stamp", sorted(k[5:] for k in GX.files if k.startswith("data_")))
def test_closed_loop_reproduces_the_reference_runs_per_seed(torch_cuda, stamp):
    """The reference's own result files (src/simulation/test_data, written by experiments.py with the real acados
    solver; tests/golden/make_experiment_golden.py) hold the outcome of 100 seeded episodes per setting.  The batched GPU
    closed loop replays the same seeds (same obstacle scenarios and noise streams).  Episodes are chaotic once a QP is cut
    off by QP_ITER, so agreement is statistical -- but it is PER SEED: observed on B200 (round 1) with QP_ITER=100: 54-63 %
    of the episodes end with exactly the reference's number of control steps, hit and reached flags (min margin equal to
    1e-3); QP_ITER=50: 35-47 %; QP_ITER=25: 20-26 %; same hit/reached outcome 76-92 %; median |steps difference| 0-8."""
    from b200mpc import experiments
    ref = GX["data_" + stamp]
    scen, tf, N, M, qp = GX["spec_" + stamp]
    ours = experiments.run_scenario("RANDOM" if scen == 0 else "EDGE", range(100), world.MpcDims(int(N), int(M), float(tf)), qp_iter_max=int(qp))
    same_outcome = ((ref[:, 0] == ours[:, 0]) & (ref[:, 1] == ours[:, 1]))
    identical = same_outcome & (ref[:, 4] == ours[:, 4])
    both = (ref[:, 1] == 1) & (ours[:, 1] == 1)
    assert same_outcome.mean() >= 0.65, same_outcome.mean()
    assert identical.mean() >= {100: 0.35, 50: 0.20, 25: 0.08}[int(qp)], identical.mean()
    assert np.median(np.abs(ref[both, 4] - ours[both, 4])) <= 20
    assert np.corrcoef(ref[both, 4], ours[both, 4])[0, 1] >= 0.45
    assert abs(ours[:, 4].mean() - ref[:, 4].mean()) <= 0.25 * ref[:, 4].mean()
    assert abs(ours[:, 0].mean() - ref[:, 0].mean()) <= 0.12 and abs(ours[:, 1].mean() - ref[:, 1].mean()) <= 0.12
