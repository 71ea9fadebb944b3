"""Vectorised RL environment with the batched GPU MPC as the low-level controller.

API shape of the reference's (non-functional) training stub, src/train.py:10-29: `reset() -> state`,
`step(action) -> (next_state, reward, done, info)`, state dimension 3*(n_obst+1), action dimension 2 =
the next sub-goal (x, y) handed to `set_subgoal` (robot_ocp_problem.py:279-284; here stage and terminal
references move together -- the reference's own set_subgoal is dimensionally broken).
Every env.step runs `mpc_steps` closed-loop control steps of robot_ocp_problem.py:184-260 on the device.
The reference ships no reward; the one here (progress towards the final goal, collision / out-of-bounds
penalties) is a placeholder and is documented as such."""
from __future__ import annotations

import numpy as np

from . import world
from .closed_loop import BatchedRobotOcp


class BatchedMpcEnv:
    def __init__(self, n_envs, dims: world.MpcDims = world.MpcDims(), mpc_steps=5, max_env_steps=80, seed=0,
                 scenario="RANDOM", random_move=True, device=0):
        self.B, self.dims, self.mpc_steps, self.max_env_steps = n_envs, dims, mpc_steps, max_env_steps
        self.scenario, self.random_move, self.device = scenario, random_move, device
        self.rng = np.random.default_rng(seed)
        self.goal = np.tile(np.array([world.X_MAX_ROBOT, world.Y_MAX_ROBOT]), (n_envs, 1))
        self.ocp = None
        self.state_dim = 3 * (dims.M + 1)
        self.action_dim = 2

    def _observe(self):
        s = self.ocp.solver
        x0 = s.get("x0", self.B, host=True); obst = s.get("obst", self.B, host=True)
        robot = np.concatenate([x0[:, :2], np.full((self.B, 1), world.R_ROBOT)], axis=1)
        obs = np.concatenate([obst[:, :, :2], np.full((self.B, self.dims.M, 1), world.R_OBST)], axis=2)
        return np.concatenate([robot, obs.reshape(self.B, -1)], axis=1), x0, obst

    def reset(self):
        B = self.B
        start = np.zeros((B, 5))
        start[:, 0] = self.rng.uniform(world.X_MIN_ROBOT, world.X_MIN_ROBOT + 2, B)
        start[:, 1] = self.rng.uniform(world.Y_MIN_ROBOT, world.Y_MIN_ROBOT + 2, B)
        start[:, 2] = np.pi / 4
        obst = world.sample_obstacles(self.rng, B, self.dims.M, self.scenario)
        if self.ocp is None:
            self.ocp = BatchedRobotOcp(start, self.goal, obst, self.dims, init_guess_when_error=True,
                                       random_move=self.random_move, device=self.device)
        else:
            self.ocp.set_up_new_experiment(obst, robot_init=start)
        self.ocp.set_initial_guess()
        self.t = 0
        self.alive = np.ones(B, bool)
        state, x0, _ = self._observe()
        self._dist = np.linalg.norm(x0[:, :2] - self.goal, axis=1)
        return state

    def step(self, action):
        """action[B,2]: sub-goal; returns (state[B,3(M+1)], reward[B], done[B], info)"""
        B, s = self.B, self.ocp.solver
        action = np.clip(np.asarray(action, np.float64).reshape(B, 2), world.X_MIN_ROBOT, world.X_MAX_ROBOT)
        self.ocp.set_subgoal(action)
        s.loop_begin(B)                                   # per-action flags and step counter; the episode's min margin is kept
        self.ocp.advance(self.mpc_steps, check_every=self.mpc_steps)
        state, x0, _ = self._observe()
        margin = s.get("margin", B, host=True)[:, 0]
        oob = s.get_int("oob", B, host=True).astype(bool)
        dist = np.linalg.norm(x0[:, :2] - self.goal, axis=1)
        hit = margin <= 0
        reached = dist <= world.TOL
        reward = (self._dist - dist) - 10.0 * hit - 10.0 * oob + 10.0 * reached
        reward = np.where(self.alive, reward, 0.0)
        self._dist = dist
        self.t += 1
        done = ~self.alive | hit | oob | reached | (self.t >= self.max_env_steps)
        self.alive &= ~done
        return state, reward, done, {"min_margin": margin, "qp_failures": s.get_int("nfail", B, host=True)}

    def close(self):
        if self.ocp is not None:
            self.ocp.solver.close()
            self.ocp = None
