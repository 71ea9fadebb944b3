"""Secondary BASELINE.json configs (not the bench.py headline): config 1 single closed-loop episode,
config 3 per-GPU share of 262144 scenarios, config 4 long horizon / many obstacles, config 5 RL env stepping.
Writes gpurun_out/configs.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200"))
import numpy as np, torch
from b200mpc import BatchedMpcSolver, world, experiments
from b200mpc.closed_loop import BatchedRobotOcp
from b200mpc.rl_env import BatchedMpcEnv
out = {}

def time_solve(N, M, tf, B, reps=3, kernel=0):
    dims = world.MpcDims(N, M, tf)
    sc = world.random_scenarios(B, dims, seed=0)
    t = {k: torch.from_numpy(sc[k]).cuda() for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")}
    s = BatchedMpcSolver(N, M, tf, max_batch=B, kernel=kernel)
    for k in ("x0", "p", "yref", "zl", "Zl"):
        s.set(k, t[k])
    ms = []
    for _ in range(reps):
        for k in ("x", "u", "sl", "su"):
            s.set(k, t[k])
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(B); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    it = s.get_int("qp_iter", B).float().mean().item()
    name = s.kernel_name
    s.close()
    return dict(N=N, M=M, tf=tf, B=B, kernel=name, ms=min(ms), solves_per_s=B / min(ms) * 1e3, mean_qp_iter=it)

# config 1: one closed-loop episode as shipped (start (-6,-6), goal (6,6), RANDOM, noisy obstacles), B = 1
obst, noise = world.reference_episode_streams(0, "RANDOM", 5, 400, True)
ocp = BatchedRobotOcp(np.array([-6.0, -6.0, np.pi / 4, 0, 0]), np.array([6.0, 6.0]), obst[None], world.MpcDims(), init_guess_when_error=True,
                      random_move=True, noise=noise[:, None])
torch.cuda.synchronize(); t0 = time.perf_counter()
res = ocp.step(400, check_every=1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
steps = int(res[5][0]) + (1 if res[2][0] else 0)
out["config1_single_episode"] = dict(control_steps=steps, reached=bool(res[2][0]), hit=bool(res[1][0]), wall_s=dt, ms_per_control_step=1e3 * dt / max(steps, 1))
ocp.solver.close()
# config 2/3: per-GPU batches of the 262144-scenario sweep (8 / 4 / 2 GPUs -> 32768 / 65536 / 131072 per GPU)
out["config3_per_gpu_share"] = [time_solve(20, 5, 2.0, B, reps=2) for B in (4096, 32768, 65536, 131072)]
# the reference's other recorded setting: TF = 1 s => N_SOLV = 10 (dedicated instantiation)
out["ref_setting_tf1_n10"] = [time_solve(10, 5, 1.0, B, reps=3) for B in (4096, 32768)]
# config 4: long horizon, many obstacles (warp kernel, v/pi rows in global memory: 3 CTAs/SM)
out["config4_long_horizon"] = [time_solve(100, 20, 5.0, B, reps=2) for B in (4096,)]
# closed loop on device: 4096 environments x 50 control steps
B = 4096
rng = np.random.default_rng(0)
ocp = BatchedRobotOcp(np.array([-6.0, -6.0, np.pi / 4, 0, 0]), np.array([6.0, 6.0]), world.sample_obstacles(rng, B), world.MpcDims(),
                      init_guess_when_error=True, random_move=True)
ocp.set_initial_guess(); ocp.advance(5, check_every=100)
torch.cuda.synchronize(); t0 = time.perf_counter(); ocp.advance(50, check_every=100); torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["closed_loop_4096_envs"] = dict(control_steps=50, wall_s=dt, env_steps_per_s=B * 50 / dt, mean_qp_iter=float(ocp.solver.get_int("qp_iter", B).float().mean()))
ocp.solver.close()
# config 5: RL loop, 65536 parallel envs, stub policy proposing random sub-goals
env = BatchedMpcEnv(65536, mpc_steps=5, max_env_steps=1000, seed=0)
env.reset(); env.step(env.goal)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    a = env.rng.uniform(-6, 6, (65536, 2))
    env.step(a)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["config5_rl_65536_envs"] = dict(env_steps=4, mpc_steps_per_env_step=5, wall_s=dt, env_steps_per_s=65536 * 4 / dt, mpc_solves_per_s=65536 * 20 / dt)
env.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
