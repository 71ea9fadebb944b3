#!/usr/bin/env python3
"""Generate tests/golden/reference_golden.npz by EXECUTING the reference's own Python code
(/root/reference, read-only) wherever that is possible without acados/CasADi binaries:

  * src/utils/visualization.py:Obstacle  (step / predict_step / predict_trajectory incl. the vx=vy quirk)
  * src/utils/obstacle_generator.py:generate_random_moving_obstacles  (scenario sampling from np.random.seed)
  * src/simulation/robot_ocp_problem.py:RobotOcpProblem  init_ocp / parameterize_slack / parameterize_model /
    set_initial_guess / shift, driven against a RECORDING stand-in for acados_template (no solver)
  * src/models/robot_model.py:export_robot_ode_model  evaluated numerically through the SymPy casadi shim

The acados solve itself (robot_ocp_problem.py:195) cannot be executed here (parity unpinned, see
oracle/mpc_oracle.h); everything AROUND it is pinned by these vectors.
Run once in the build container:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
SHIMS = os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200", "codegen", "shims")

# ---- stand-ins for what the reference imports but this image lacks -------------------------------
mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot"); mpl.animation = types.ModuleType("matplotlib.animation")
sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot, "matplotlib.animation": mpl.animation})
sys.path.insert(0, SHIMS)                                  # casadi (SymPy), imp, acados_template.AcadosModel
import acados_template  # noqa: E402

LOG = []


class _Bag:
    pass


class AcadosOcp:
    def __init__(self):
        self.model = None; self.dims = _Bag(); self.cost = _Bag(); self.constraints = _Bag(); self.solver_options = _Bag()
        self.parameter_values = None


class AcadosOcpSolver:
    def __init__(self, ocp, json_file=None):
        self.ocp = ocp; self.store = {}

    def set(self, stage, field, value): LOG.append(("set", stage, field, np.array(value, float).copy())); self.store[(stage, field)] = np.array(value, float).copy()
    def cost_set(self, stage, field, value): LOG.append(("cost_set", stage, field, np.array(value, float).copy()))
    def set_params_sparse(self, stage, idx, value): LOG.append(("set_params_sparse", stage, np.array(idx).copy(), np.array(value, float).copy()))
    def get(self, stage, field): return self.store[(stage, field)].copy()
    def reset(self): LOG.append(("reset",))
    def solve(self): raise RuntimeError("the acados solve cannot be executed in this image")


class AcadosSimSolver:
    def __init__(self, ocp, json_file=None): pass


for name, cls in (("acados_ocp", AcadosOcp), ("acados_ocp_solver", AcadosOcpSolver), ("acados_sim_solver", AcadosSimSolver)):
    m = types.ModuleType("acados_template." + name); setattr(m, cls.__name__, cls); sys.modules["acados_template." + name] = m
    setattr(acados_template, name, m)
acados_template.AcadosOcp = AcadosOcp; acados_template.AcadosOcpSolver = AcadosOcpSolver; acados_template.AcadosSimSolver = AcadosSimSolver

sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(REF, "src", "simulation"))
from models import world_specification as ws                      # noqa: E402  (reference, unmodified)
from utils.visualization import Obstacle                          # noqa: E402
from utils.obstacle_generator import generate_random_moving_obstacles   # noqa: E402
from models.robot_model import export_robot_ode_model             # noqa: E402
import sympy as sp                                                # noqa: E402

out = {}
assert abs(ws.TF / ws.N_SOLV - 0.1) < 1e-15
out["world_constants"] = np.array([ws.X_MIN, ws.X_MAX, ws.Y_MIN, ws.Y_MAX, ws.R_ROBOT, ws.V_MAX_ROBOT, ws.C_MAX, ws.N_OBST, ws.R_OBST,
                                   ws.RANDOMNESS, ws.V_MAX_OBST, ws.MARGIN, ws.X_MIN_OBST, ws.X_MAX_OBST, ws.Y_MIN_OBST, ws.Y_MAX_OBST,
                                   ws.TOL, ws.QP_ITER, ws.X_MIN_ROBOT, ws.X_MAX_ROBOT], float)

# ---- 1. scenario sampling + obstacle motion + horizon prediction ---------------------------------
import io, contextlib  # noqa: E402
STEPS, NPRED = 40, 20
for scen in ("RANDOM", "EDGE", "CENTER"):
    init, traj, pred0, pred_mid = [], [], [], []
    for seed in range(6):
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            obs = generate_random_moving_obstacles(scen, True)
        init.append([[o.x, o.y, o.vx, o.vy] for o in obs])
        pred0.append([o.predict_trajectory(NPRED) for o in obs])
        tr = []
        for k in range(STEPS):
            for o in obs:
                o.step()
            tr.append([[o.x, o.y, o.vx, o.vy] for o in obs])
            if k == STEPS // 2:
                pred_mid.append([o.predict_trajectory(NPRED) for o in obs])
        traj.append(tr)
    out[f"obst_init_{scen}"] = np.array(init, float)          # [seed, M, 4]
    out[f"obst_traj_{scen}"] = np.array(traj, float)          # [seed, step, M, 4]
    out[f"obst_pred0_{scen}"] = np.array(pred0, float)        # [seed, M, NPRED+1, 2]
    out[f"obst_predmid_{scen}"] = np.array(pred_mid, float)

# ---- 2. RobotOcpProblem against the recording solver ----------------------------------------------
from robot_ocp_problem import RobotOcpProblem                    # noqa: E402
np.random.seed(3)
x_init = np.array([ws.X_MIN + 2, ws.Y_MIN + 2, np.pi / 4, 0.7, -0.3])
goal = np.array([ws.X_MAX - 2, ws.Y_MAX - 2])
with contextlib.redirect_stdout(io.StringIO()):
    prob = RobotOcpProblem(x_init.copy(), goal.copy(), scenario="RANDOM", slack=True, random_move=True)
ocp = prob.ocp
out["ocp_N"] = np.array([ocp.dims.N])
for k in ("Vx", "Vu", "yref", "Vx_e", "yref_e", "W", "W_e", "zl", "Zl", "zu", "Zu"):
    out["cost_" + k] = np.array(getattr(ocp.cost, k), float)
for k in ("x0", "lbx", "ubx", "idxbx", "lbu", "ubu", "idxbu", "lh", "uh", "lh_e", "uh_e", "Jsh", "Jsh_e"):
    out["con_" + k] = np.array(getattr(ocp.constraints, k), float)
out["opt_lm_tf_qpiter"] = np.array([ocp.solver_options.levenberg_marquardt, ocp.solver_options.tf, ocp.solver_options.qp_solver_iter_max], float)
out["opt_strings"] = np.array([ocp.solver_options.qp_solver, ocp.solver_options.hessian_approx, ocp.solver_options.integrator_type,
                               ocp.solver_options.nlp_solver_type, ocp.cost.cost_type, ocp.cost.cost_type_e])
# parameterize_slack (:145-152) at a moving state
prob.x0 = np.array([-3.0, 1.5, 0.2, 2.5, -0.4]); del LOG[:]
prob.parameterize_slack()
out["slack_x0"] = prob.x0.copy(); out["slack_goal"] = prob.subgoal.copy()
out["slack_zl"] = np.array([v for (c, s, f, v) in [e for e in LOG if e[0] == "cost_set"] if f == "zl"])
out["slack_Zl"] = np.array([v for (c, s, f, v) in [e for e in LOG if e[0] == "cost_set"] if f == "Zl"])
# parameterize_model (:154-166)
del LOG[:]
prob.parameterize_model()
out["model_obst"] = np.array([[o.x, o.y, o.vx, o.vy] for o in prob.obstacles], float)
out["model_p"] = np.array([e[3] for e in LOG if e[0] == "set" and e[2] == "p"])
# set_initial_guess (:286-306) incl. the aliasing quirk
prob.x0 = np.array([1.0, -2.0, 0.3, 4.0, 5.0]); del LOG[:]
prob.set_initial_guess()
out["guess_x0_after"] = prob.x0.copy()
out["guess_x"] = np.array([e[3] for e in LOG if e[0] == "set" and e[2] == "x"])
out["guess_u"] = np.array([e[3] for e in LOG if e[0] == "set" and e[2] == "u"])
# warm-start shift (:253-258) replayed with the reference's own statements on the recording solver
N = ocp.dims.N
rng = np.random.RandomState(0)
X = rng.randn(N + 1, 5); U = rng.randn(N, 2)
for i in range(N + 1):
    prob.ocp_solver.set(i, "x", X[i])
for i in range(N):
    prob.ocp_solver.set(i, "u", U[i])
for j in range(N - 1):
    prob.ocp_solver.set(j, 'x', prob.ocp_solver.get(j + 1, 'x'))
    prob.ocp_solver.set(j, 'u', prob.ocp_solver.get(j + 1, 'u'))
prob.ocp_solver.set(N - 1, 'x', prob.ocp_solver.get(N, 'x'))
prob.ocp_solver.set(N - 1, 'u', np.array([0, 0]))
out["shift_X_in"] = X; out["shift_U_in"] = U
out["shift_X_out"] = np.array([prob.ocp_solver.get(i, "x") for i in range(N + 1)])
out["shift_U_out"] = np.array([prob.ocp_solver.get(i, "u") for i in range(N)])

# ---- 3. model expressions evaluated numerically ---------------------------------------------------
model = export_robot_ode_model()
xs, us, ps = model.x.elems, model.u.elems, model.p.elems
f = sp.lambdify([xs, us], model.f_expl_expr.elems, "numpy")
h = sp.lambdify([xs, ps], model.con_h_expr.elems, "numpy")
rng = np.random.RandomState(1)
XX = np.column_stack([rng.uniform(-7, 7, 50), rng.uniform(-7, 7, 50), rng.uniform(-np.pi, np.pi, 50), rng.uniform(-10, 10, 50), rng.uniform(-10, 10, 50)])
UU = rng.uniform(-8, 8, (50, 2)); PP = rng.uniform(-6, 6, (50, 2 * ws.N_OBST))
out["model_x"] = XX; out["model_u"] = UU; out["model_pp"] = PP
out["model_f"] = np.array([f(x, u) for x, u in zip(XX, UU)], float)
out["model_h"] = np.array([h(x, p) for x, p in zip(XX, PP)], float)
out["model_dims"] = np.array([len(xs), len(us), len(ps)])

np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
print("wrote", os.path.join(HERE, "reference_golden.npz"), {k: v.shape for k, v in out.items()})
