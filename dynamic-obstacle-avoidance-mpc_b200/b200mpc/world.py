"""World constants and scenario sampling for the batched solver.

Mirrors src/models/world_specification.py:7-48 (constants),
src/utils/obstacle_generator.py:8-28 (scenario distribution) and the deterministic part
of src/utils/visualization.py:25-79 (constant-velocity obstacles with wall bounce, horizon
prediction incl. the `vx = self.vy` quirk at :69) as vectorised numpy over a batch of
independent environments.  Host-side input generation only; no solver arithmetic here.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# world_specification.py:7-48
X_MIN, X_MAX, Y_MIN, Y_MAX = -8.0, 8.0, -8.0, 8.0
R_ROBOT = 0.2
V_MAX_ROBOT = 10.0
X_MIN_ROBOT, X_MAX_ROBOT, Y_MIN_ROBOT, Y_MAX_ROBOT = -6.0, 6.0, -6.0, 6.0
C_MAX = 8.0
N_OBST = 5
R_OBST = 1.0
RANDOMNESS = 0.1
V_MAX_OBST = 2.0
MARGIN = 1.2
Y_MIN_OBST = Y_MIN_ROBOT + 1.0 + 3 * R_ROBOT      # = -4.4
Y_MAX_OBST = -Y_MIN_ROBOT                          # = 6
X_MIN_OBST, X_MAX_OBST = Y_MIN_OBST, Y_MAX_ROBOT
TOL = 0.15
QP_ITER = 50


@dataclasses.dataclass(frozen=True)
class MpcDims:
    """Horizon / obstacle count / horizon length (TF, N_SOLV, N_OBST of the reference)."""
    N: int = 20
    M: int = N_OBST
    tf: float = 2.0

    @property
    def dt(self) -> float:
        return self.tf / self.N


def predict_step(x, vx, y, vy, dt):
    """visualization.py:35-59, vectorised (noise handled by the caller)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        thx = np.where(vx < 0, (x - X_MIN) / np.abs(vx), np.where(vx > 0, (X_MAX - x) / np.abs(vx), np.inf))
        thy = np.where(vy < 0, (y - Y_MIN) / np.abs(vy), np.where(vy > 0, (Y_MAX - y) / np.abs(vy), np.inf))
    hx = thx <= dt
    x2 = np.where(hx, x + (vx * thx - vx * (dt - thx)), x + vx * dt)
    vx2 = np.where(hx, -vx, vx)
    hy = thy <= dt
    y2 = np.where(hy, y + (vy * thy - vy * (dt - thy)), y + vy * dt)
    vy2 = np.where(hy, -vy, vy)
    return x2, vx2, y2, vy2


def obstacle_step(obst, dt, random_move=False, noise=None):
    """visualization.py:20-33: obst[...,4] = (x,y,vx,vy); noise[...,2] ~ N(0,1)."""
    x, y, vx, vy = (obst[..., i] for i in range(4))
    if random_move:
        vx = np.clip((1 + RANDOMNESS * noise[..., 0]) * vx, -V_MAX_OBST, V_MAX_OBST)
        vy = np.clip((1 + RANDOMNESS * noise[..., 1]) * vy, -V_MAX_OBST, V_MAX_OBST)
    x, vx, y, vy = predict_step(x, vx, y, vy, dt)
    return np.stack([x, y, vx, vy], axis=-1)


def predict_trajectory(obst, n, dt, quirk=True):
    """visualization.py:62-79 -> P[..., n+1, M, 2]; quirk reproduces `vx = self.vy` (:69)."""
    x, y, vy = obst[..., 0], obst[..., 1], obst[..., 3]
    vx = obst[..., 3] if quirk else obst[..., 2]
    out = np.zeros(obst.shape[:-2] + (n + 1,) + obst.shape[-2:-1] + (2,))
    out[..., 0, :, 0] = x
    out[..., 0, :, 1] = y
    for i in range(n):
        x, vx, y, vy = predict_step(x, vx, y, vy, dt)
        out[..., i + 1, :, 0] = x
        out[..., i + 1, :, 1] = y
    return out


def sample_obstacles(rng, B, M=N_OBST, scenario="RANDOM"):
    """obstacle_generator.py:8-28 -> obst[B,M,4]."""
    if scenario == "RANDOM":
        xs = rng.uniform(X_MIN_OBST, X_MAX_OBST, (B, M))
        ys = rng.uniform(Y_MIN_OBST, Y_MAX_OBST, (B, M))
    elif scenario == "CENTER":
        xs = np.zeros((B, M)); ys = np.zeros((B, M))
    elif scenario == "EDGE":
        xs = 7 * np.ones((B, M)); ys = 7 * np.ones((B, M))
    else:
        raise ValueError(f"unknown scenario {scenario!r}")
    vx = rng.uniform(-V_MAX_OBST, V_MAX_OBST, (B, M))
    vy = rng.uniform(-V_MAX_OBST, V_MAX_OBST, (B, M))
    return np.stack([xs, ys, vx, vy], axis=-1)


def slack_weights(x0, goal, N, M):
    """robot_ocp_problem.py:145-152 -> zl, Zl [B,N+1,M] (equal)."""
    d = (x0[:, 0] - goal[:, 0]) ** 2 + (x0[:, 1] - goal[:, 1]) ** 2 + x0[:, 3] ** 2 + x0[:, 4] ** 2
    scale = 1e4 * (d + 50.0)
    alpha = scale[:, None] * ((N - np.arange(N + 1)) / N)[None, :]
    zl = np.repeat(alpha[:, :, None], M, axis=2)
    return zl, zl.copy()


def random_scenarios(B, dims: MpcDims = MpcDims(), seed=0, scenario="RANDOM", moving_start=False):
    """Synthetic benchmark input: random start pose in the robot box,
    goal (6,6), RANDOM obstacles; warm iterate = set_initial_guess (robot_ocp_problem.py:286-306).
    With moving_start the robot also gets a random v, omega (exercises dx0 != 0 paths)."""
    rng = np.random.default_rng(seed)
    N, M = dims.N, dims.M
    x0 = np.zeros((B, 5))
    x0[:, 0] = rng.uniform(X_MIN_ROBOT, X_MAX_ROBOT, B)
    x0[:, 1] = rng.uniform(Y_MIN_ROBOT, Y_MAX_ROBOT, B)
    x0[:, 2] = rng.uniform(-np.pi, np.pi, B)
    if moving_start:
        x0[:, 3] = rng.uniform(-3, 3, B)
        x0[:, 4] = rng.uniform(-1, 1, B)
    goal = np.tile(np.array([X_MAX_ROBOT, Y_MAX_ROBOT]), (B, 1))
    obst = sample_obstacles(rng, B, M, scenario)
    P = predict_trajectory(obst, N, dims.dt)                    # [B,N+1,M,2]
    p = np.ascontiguousarray(P.reshape(B, N + 1, 2 * M))
    yref = np.zeros((B, N + 1, 6))
    yref[:, :, 0:2] = goal[:, None, :]
    zl, Zl = slack_weights(x0, goal, N, M)
    x = np.repeat(x0[:, None, :], N + 1, axis=1).copy()
    x[:, :, 3:] = 0.0
    u = np.zeros((B, N, 2))
    sl = np.zeros((B, N + 1, M)); su = np.zeros((B, N + 1, M))
    return dict(x0=x0, goal=goal, obst=obst, p=p, yref=yref, zl=zl, Zl=Zl, x=x, u=u, sl=sl, su=su)


def reference_episode_streams(seed, scenario="RANDOM", M=N_OBST, steps=400, random_move=True):
    """The obstacle initial states and the velocity-noise stream ONE reference episode consumes after
    `np.random.seed(seed)` (experiments.py:33-36): generate_random_moving_obstacles draws x, y (RANDOM
    only), vx, vy as (M,1) uniforms (obstacle_generator.py:10-22); every control step then draws
    `np.random.normal(size=2)` per obstacle in order (visualization.py:31).  Returns obst[M,4],
    noise[steps,M,2] -- bit-identical to what the reference's global legacy RNG would produce."""
    rs = np.random.RandomState(seed)
    if scenario == "RANDOM":
        xs = rs.uniform(X_MIN_OBST, X_MAX_OBST, (M, 1)); ys = rs.uniform(Y_MIN_OBST, Y_MAX_OBST, (M, 1))
    elif scenario == "CENTER":
        xs = np.zeros((M, 1)); ys = np.zeros((M, 1))
    elif scenario == "EDGE":
        xs = 7 * np.ones((M, 1)); ys = 7 * np.ones((M, 1))
    else:
        raise ValueError(scenario)
    vx = rs.uniform(-V_MAX_OBST, V_MAX_OBST, (M, 1)); vy = rs.uniform(-V_MAX_OBST, V_MAX_OBST, (M, 1))
    obst = np.hstack((xs, ys, vx, vy))
    noise = rs.normal(size=(steps, M, 2)) if random_move else np.zeros((steps, M, 2))
    return obst, noise
