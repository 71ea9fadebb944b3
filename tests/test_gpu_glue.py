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


# ---------------------------------------------------------------------------------------------------------------------
# round 2: fused closed-loop entry (b2mpc_closed_loop), device Philox noise, re-entrant step(), RL env against the oracle
def _philox4x32_10(c, k):
    """numpy restatement of Philox4x32-10 (Salmon et al., SC'11) for one counter/key: the check of the device generator"""
    c = [np.uint32(v) for v in c]; k = [np.uint32(v) for v in k]
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0 = M0 * np.uint64(c[0]); p1 = M1 * np.uint64(c[2])
        c = [np.uint32(p1 >> np.uint64(32)) ^ c[1] ^ k[0], np.uint32(p1 & np.uint64(0xffffffff)),
             np.uint32(p0 >> np.uint64(32)) ^ c[3] ^ k[1], np.uint32(p0 & np.uint64(0xffffffff))]
        k = [np.uint32((int(k[0]) + 0x9E3779B9) & 0xffffffff), np.uint32((int(k[1]) + 0xBB67AE85) & 0xffffffff)]
    return [int(v) for v in c]


def _philox_normal2(seed, env, step, obstacle):
    r = _philox4x32_10([step, obstacle, env & 0xffffffff, env >> 32], [seed & 0xffffffff, seed >> 32])
    u1 = ((r[0] >> 5) * 67108864.0 + (r[1] >> 6) + 0.5) / 9007199254740992.0
    u2 = ((r[2] >> 5) * 67108864.0 + (r[3] >> 6) + 0.5) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)


def test_philox_known_answer():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors): zero counter/key, all-ones, and pi digits"""
    assert _philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_device_philox_noise_stream(torch_cuda):
    """noise_mode 1 of b2mpc_closed_loop: the obstacle velocities after each control step reveal the N(0,1) draws
    (visualization.py:28-33: v <- clip((1 + 0.1 n) v)); they must be the Philox4x32-10 + Box-Muller stream keyed by
    (seed, global environment index, control step, obstacle) -- for any batch composition."""
    from b200mpc.closed_loop import BatchedRobotOcp
    dims = world.MpcDims(20, 5, 2.0)
    B, seed = 6, 0x1234567
    obst = np.zeros((B, 5, 4)); obst[:, :, 0] = 3.0; obst[:, :, 1] = -3.0; obst[:, :, 2] = 1.0; obst[:, :, 3] = -1.0
    start = np.array([-6.0, -6.0, np.pi / 4, 0.0, 0.0]); goal = np.array([6.0, 6.0])
    ocp = BatchedRobotOcp(start, goal, obst, dims, init_guess_when_error=True, random_move=True, seed=seed)
    ocp.set_initial_guess()
    v_prev = obst[:, :, 2:4].copy()
    for step in range(3):
        ocp.advance(1)
        v = ocp.solver.get("obst", B, host=True)[:, :, 2:4]
        n = (v / v_prev - 1.0) / 0.1
        for b in range(B):
            for j in range(5):
                n0, n1 = _philox_normal2(seed, b, step, j)
                assert abs(n[b, j, 0] - n0) < 1e-9 and abs(n[b, j, 1] - n1) < 1e-9, (step, b, j)
        v_prev = v.copy()
    ocp.solver.close()
    # environments 3..5 alone, told their global index: the same stream, the same trajectories
    sub = BatchedRobotOcp(start, goal, obst[3:], dims, init_guess_when_error=True, random_move=True, seed=seed, env_offset=3)
    sub.set_initial_guess(); sub.advance(3)
    full = BatchedRobotOcp(start, goal, obst, dims, init_guess_when_error=True, random_move=True, seed=seed)
    full.set_initial_guess(); full.advance(3)
    assert np.array_equal(sub.solver.get("obst", 3, host=True), full.solver.get("obst", B, host=True)[3:])
    assert np.array_equal(sub.solver.get("x0", 3, host=True), full.solver.get("x0", B, host=True)[3:])
    sub.solver.close(); full.solver.close()


def test_fused_closed_loop_equals_the_stepwise_loop(torch_cuda):
    """b2mpc_closed_loop (one C call per chunk, three launches per control step, finished environments leave the queue)
    against the loop sequenced phase by phase from Python: bit-identical states, margins, flags and step counts"""
    from b200mpc.closed_loop import BatchedRobotOcp
    dims = world.MpcDims(20, 5, 2.0)
    seeds, steps = list(range(24)), 150
    streams = [world.reference_episode_streams(s, "RANDOM", 5, steps, True) for s in seeds]
    obst = np.stack([o for o, _ in streams]); noise = np.ascontiguousarray(np.stack([n for _, n in streams], axis=1))
    start = np.array([-7.0, -7.0, np.pi / 4, 0.0, 0.0]); goal = np.array([7.0, 7.0])
    res = []
    for fused in (True, False):
        ocp = BatchedRobotOcp(start, goal, obst, dims, init_guess_when_error=True, random_move=True, noise=noise)
        ocp.solver.loop_begin(len(seeds)); ocp.set_initial_guess()
        (ocp.advance if fused else ocp.advance_stepwise)(steps, check_every=16)
        res.append(ocp.outcome() + (ocp.solver.get("x", len(seeds), host=True), ocp.solver.get("obst", len(seeds), host=True)))
        ocp.solver.close()
    assert res[1][2].sum() >= 1, (res[0][2].sum(), res[1][2].sum(), res[0][5], res[1][5])    # some environments reached the goal (left the queue)
    running = ~res[0][2]
    for i, (a, b) in enumerate(zip(*res)):
        # the warm-start iterate (index 7) of an environment that has reached its goal is frozen by the fused loop (it leaves
        # the queue) while the stepwise loop keeps solving it: compared for the environments that still run
        assert np.array_equal(a[running], b[running]) if i == 7 else np.array_equal(a, b), i


def test_step_is_reentrant_after_set_subgoal(torch_cuda):
    """RobotOcpProblem.step() starts every call with reached_subgoal = False and i = 0 (robot_ocp_problem.py:177-183): after an
    environment reached its sub-goal, set_subgoal(new) + step() must drive it on; min_margin_traj is kept (:51)"""
    from b200mpc.closed_loop import BatchedRobotOcp
    dims = world.MpcDims(20, 5, 2.0)
    B = 4
    obst = np.zeros((B, 5, 4)); obst[:, :, 0] = 7.5; obst[:, :, 1] = -7.5          # parked in a corner, out of the way
    start = np.array([-5.0, -5.0, np.pi / 4, 0.0, 0.0])
    ocp = BatchedRobotOcp(start, np.array([-3.0, -3.0]), obst, dims, init_guess_when_error=True, random_move=False)
    x1, hit1, reached1, mm1, dist1, steps1, oob1 = ocp.step(120)
    assert reached1.all() and (steps1 > 3).all() and (dist1 <= world.TOL).all()
    ocp.set_subgoal(np.array([0.0, -3.0]))
    x2, hit2, reached2, mm2, dist2, steps2, oob2 = ocp.step(120)
    assert reached2.all() and (np.abs(x2[:, 0] - 0.0) < 0.2).all()          # it moved to the NEW sub-goal
    assert (steps2 > 3).all() and (steps2 < 120).all()                      # per-call step count, not cumulative
    assert (mm2 <= mm1 + 1e-12).all()                                       # the episode's minimum margin is kept
    ocp.solver.close()


def test_set_subgoal_moves_stage_and_terminal_references(torch_cuda):
    """robot_ocp_problem.py:279-284, fixed variant (the reference writes a 5-vector into the 4-dimensional terminal
    reference and leaves the stage references at the old goal): yref of EVERY node and the slack-weight goal follow"""
    from b200mpc.closed_loop import BatchedRobotOcp
    dims = world.MpcDims(20, 5, 2.0)
    B = 3
    obst = np.zeros((B, 5, 4)); obst[:, :, 0] = 7.5; obst[:, :, 1] = 7.5
    ocp = BatchedRobotOcp(np.array([0.0, 0.0, 0.0, 0.0, 0.0]), np.array([6.0, 6.0]), obst, dims)
    sub = np.array([[1.0, 2.0], [-3.0, 4.0], [5.0, -6.0]])
    ocp.set_subgoal(sub)
    yref = ocp.solver.get("yref", B, host=True)
    assert yref.shape == (B, 21, 6)
    assert np.array_equal(yref[:, :, 0:2], np.repeat(sub[:, None, :], 21, axis=1)) and (yref[:, :, 2:] == 0).all()
    assert np.array_equal(ocp.solver.get("goal", B, host=True), sub)
    ocp.solver.prepare_slack(B)
    zl = ocp.solver.get("zl", B, host=True)
    exp = 1e4 * ((sub ** 2).sum(1) + 50.0)                                  # x0 = 0: alpha_0 = 1e4 (|x0 - subgoal|^2 + 50)
    assert np.allclose(zl[:, 0, 0], exp, rtol=1e-15)
    ocp.solver.close()


def test_rl_env_step_tracks_the_oracle_loop(torch_cuda, oracle):
    """BatchedMpcEnv.step(action) = set_subgoal(action) + mpc_steps control steps of robot_ocp_problem.py:184-260: the robot
    states it returns are those of the oracle-driven loop with the same sub-goal (compared while the QPs converge)"""
    from b200mpc.rl_env import BatchedMpcEnv
    B, mpc_steps = 8, 5
    env = BatchedMpcEnv(B, mpc_steps=mpc_steps, max_env_steps=30, seed=11, random_move=False)
    state = env.reset()
    _, x0, obst0 = env._observe()
    action = np.stack([np.linspace(-2, 3, B), np.linspace(1, -4, B)], axis=1)
    cfg = oracle.default_config(20, 5, 2.0)
    checked = 0
    state1, reward, done, info = env.step(action)
    it_gpu = env.ocp.solver.get_int("qp_iter", B, host=True)
    for b in range(B):
        log = _oracle_episode(oracle, cfg, obst0[b], np.zeros((mpc_steps, 5, 2)), x0[b], action[b], mpc_steps)
        if any(entry[2] >= 50 for entry in log) or it_gpu[b] >= 50:
            continue                                   # a QP cut off by the iteration limit: the two loops may separate
        checked += 1
        assert np.abs(state1[b, :2] - log[-1][0][:2]).max() < 1e-5, b
        assert abs(info["min_margin"][b] - log[-1][3]) < 1e-5
    assert checked >= 3
    env.close()


def test_two_gpus_give_bit_identical_shards(torch_cuda):
    """per-problem results do not depend on the number of shards -- one handle on device 0 solving all scenarios
    against two handles (devices 0 and 1, used from ONE thread: every entry point switches to its handle's device) solving
    half each.  Skipped on a single-GPU box."""
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from b200mpc import BatchedMpcSolver, sharding
    dims = world.MpcDims(20, 5, 2.0)
    B = 512
    sc = world.random_scenarios(B, dims, seed=13, moving_start=True)
    names = ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")

    def run(device, lo, hi):
        s = BatchedMpcSolver(20, 5, 2.0, max_batch=hi - lo, device=device, kernel=2)
        with torch.cuda.device(device):
            for nme in names:
                s.set(nme, torch.from_numpy(np.ascontiguousarray(sc[nme][lo:hi])).to(f"cuda:{device}"))
            s.solve(hi - lo)
            out = (s.get("x", hi - lo).cpu().numpy(), s.get("u", hi - lo).cpu().numpy(), s.get_int("qp_iter", hi - lo).cpu().numpy())
        return s, out
    s_all, full = run(0, 0, B)
    (lo0, hi0), (lo1, hi1) = [sharding.shard_bounds(B, r, 2) for r in (0, 1)]
    sa, a = run(0, lo0, hi0)
    sb, b = run(1, lo1, hi1)                                 # created while device 0's handles are alive
    sa.solve(hi0 - lo0)                                      # ... and device 0 still works afterwards, from the same thread
    for k in range(3):
        assert np.array_equal(np.concatenate([a[k], b[k]]), full[k])
    for s in (s_all, sa, sb):
        s.close()
