"""b200mpc -- batched FP64 SQP-RTI nonlinear MPC for dynamic obstacle avoidance on B200 (sm_100a).

Hot path only: the acados solve the reference issues at src/simulation/robot_ocp_problem.py:195 and
the per-step glue around it, for thousands of independent problems per launch."""
from . import capi, world  # noqa: F401
from .solver import BatchedMpcSolver, dfma_peak_tflops, latency_kernel  # noqa: F401
