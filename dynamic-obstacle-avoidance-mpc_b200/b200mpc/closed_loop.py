"""Batched closed loop: the B200 counterpart of `RobotOcpProblem` (src/simulation/robot_ocp_problem.py:13)
for B independent environments.  Same public surface -- constructor arguments, `step(max_iter)`,
`set_subgoal`, `set_initial_guess`, `set_up_new_experiment` -- every per-environment quantity gains a
leading batch dimension.  All arithmetic runs in libb200mpc (solve + on-device glue kernels); this
class only sequences the launches the way the reference's `while i < max_iter` loop (:184-260) does."""
from __future__ import annotations

import numpy as np

from . import world
from .solver import BatchedMpcSolver


class BatchedRobotOcp:
    def __init__(self, robot_init, robot_end, obstacles, dims: world.MpcDims = world.MpcDims(), slack=True,
                 init_guess_when_error=False, random_move=False, noise=None, device=0, quirk_prediction=True,
                 qp_iter_max=world.QP_ITER, seed=0, env_offset=0, **solver_overrides):
        """robot_init [B,5] (or [5]), robot_end [B,2] (or [2]), obstacles [B,M,4] (x,y,vx,vy);
        noise: optional [steps,B,M,2] N(0,1) stream to replay (reference-identical streams:
        world.reference_episode_streams); without it, random_move draws the on-device Philox stream keyed by
        (seed, env_offset + environment, control step, obstacle) -- the same for any batch size or sharding."""
        import torch
        self.torch = torch
        obstacles = np.asarray(obstacles, np.float64)
        B, M = obstacles.shape[0], obstacles.shape[1]
        assert M == dims.M
        if not slack:
            raise NotImplementedError("slack=False (hard obstacle constraints) is not part of the accelerated path")
        self.B, self.dims = B, dims
        self.robot_init = np.broadcast_to(np.asarray(robot_init, np.float64), (B, 5)).copy()
        self.robot_end = np.broadcast_to(np.asarray(robot_end, np.float64), (B, 2)).copy()
        self.init_guess_when_error = init_guess_when_error
        self.random_move = random_move
        self.quirk_prediction = quirk_prediction
        self.seed, self.env_offset = seed, env_offset
        self.solver = BatchedMpcSolver(dims.N, dims.M, dims.tf, max_batch=B, device=device,
                                       qp_iter_max=qp_iter_max, **solver_overrides)
        self.dev = torch.device("cuda", device)
        self._noise = None if noise is None else torch.from_numpy(np.ascontiguousarray(noise, np.float64)).to(self.dev)
        self.init_experiment(obstacles)
        self.set_subgoal(self.robot_end)

    # robot_ocp_problem.py:35-52
    def init_experiment(self, obstacles):
        s = self.solver
        s.set("obst", np.ascontiguousarray(obstacles, np.float64))
        s.set("x0", self.robot_init)
        s.loop_reset(self.B)
        self._step_count = 0

    # robot_ocp_problem.py:279-284 (the reference only touches the terminal reference and passes a
    # 5-vector to a 4-dimensional yref_e; here stage and terminal references move together)
    def set_subgoal(self, goal_xy):
        goal = np.broadcast_to(np.asarray(goal_xy, np.float64), (self.B, 2)).copy()
        self.subgoal = goal
        yref = np.zeros((self.B, self.dims.N + 1, 6))
        yref[:, :, 0:2] = goal[:, None, :]
        self.solver.set("yref", yref)
        self.solver.set("goal", goal)

    # robot_ocp_problem.py:286-306
    def set_initial_guess(self, mask=None):
        self.solver.reset(self.B, mask)

    # robot_ocp_problem.py:309-311
    def set_up_new_experiment(self, obstacles, robot_init=None, noise=None):
        if robot_init is not None:
            self.robot_init = np.broadcast_to(np.asarray(robot_init, np.float64), (self.B, 5)).copy()
        if noise is not None:
            self._noise = self.torch.from_numpy(np.ascontiguousarray(noise, np.float64)).to(self.dev)
        self.init_experiment(np.asarray(obstacles, np.float64))

    # robot_ocp_problem.py:168-277
    def step(self, max_iter, check_every=8):
        """Run the RTI closed loop for at most max_iter control steps per environment.
        Returns (x_final[B,5], hit[B], reached[B], min_margin[B], dist_to_goal[B], steps[B], out_of_bounds[B])."""
        self.solver.loop_begin(self.B)          # reached_subgoal = False, out_of_bounds = False, i = 0 (:177-183)
        self.set_initial_guess()
        self.advance(max_iter, check_every)
        return self.outcome()

    def advance(self, n_steps, check_every=8):
        """the `while i < max_iter` loop body (:184-260) n_steps times, without re-initialising the guess: chunks of
        `check_every` control steps, each enqueued by ONE call into the library (b2mpc_closed_loop: three launches per control
        step, no host code in between); the host only looks at the done flags between chunks"""
        s, B = self.solver, self.B
        taken = 0
        while taken < n_steps:
            n = min(check_every, n_steps - taken)
            noise = None
            if self.random_move and self._noise is not None:
                noise = self._noise[self._step_count:self._step_count + n]
            s.closed_loop(B, n, noise=noise, seed=self.seed, env_offset=self.env_offset, random_move=self.random_move,
                          init_guess_when_error=self.init_guess_when_error, quirk=self.quirk_prediction, tol=world.TOL)
            self._step_count += n
            taken += n
            if taken < n_steps and bool(s.get_int("done", B).all()):
                break

    def advance_stepwise(self, n_steps, check_every=8):
        """the same loop sequenced from Python, one library call per phase (six launches per control step): what
        `advance` did before b2mpc_closed_loop existed; kept for tests that look between the phases"""
        s, B, torch = self.solver, self.B, self.torch
        for i in range(n_steps):
            s.obst_predict(B, quirk=self.quirk_prediction)          # parameterize_model(), :186
            s.prepare_slack(B)                                       # parameterize_slack(), :188
            s.solve(B)                                               # :191-195 (x0 embedding is the 'x0' field)
            noise = None
            if self.random_move:
                assert self._noise is not None, "advance_stepwise replays a given noise stream"
                noise = self._noise[self._step_count]
            s.loop_post(B, noise, self.random_move, self.init_guess_when_error, world.TOL)   # :198-260
            self._step_count += 1
            if (i + 1) % check_every == 0 and bool(s.get_int("done", B).all()):
                break

    def outcome(self):
        s, B = self.solver, self.B
        margin = s.get("margin", B, host=True)
        x0 = s.get("x0", B, host=True)
        return (x0, margin[:, 0] <= 0, s.get_int("done", B, host=True).astype(bool), margin[:, 0], margin[:, 1],
                s.get_int("steps", B, host=True), s.get_int("oob", B, host=True).astype(bool))
