"""The acados_template-compatible facade.  CPU part: the reference's RobotOcpProblem is
constructed and parameterised UNMODIFIED through the facade (needs /root/reference, skipped on the GPU
box) and the mirrored solver state is compared with the golden vectors; error behaviour of the facade.
GPU part: the facade's solve()/sim equals the batched solver on the same problem."""
import contextlib
import io
import os
import sys
import types

import numpy as np
import pytest

from b200mpc import acados_compat, world

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def _reference_ocp(N=20, tf=2.0):
    """the AcadosOcp the reference builds in init_ocp/init_ocp_solver (robot_ocp_problem.py:54-132), restated
    on the facade's classes so that this also runs where /root/reference is absent"""
    acados_compat.install()
    from casadi import SX, vertcat
    ocp = acados_compat.AcadosOcp()
    m = ocp.model
    m.x = vertcat(*[SX.sym(n) for n in ("x", "y", "psi", "v", "omega")]); m.u = vertcat(SX.sym("u_a"), SX.sym("u_alpha"))
    m.p = SX.sym("p", 10)
    m.con_h_expr = vertcat(*[(m.x[0] - m.p[2 * i]) ** 2 + (m.x[1] - m.p[2 * i + 1]) ** 2 - 5.76 for i in range(5)])
    m.con_h_expr_e = m.con_h_expr
    ocp.dims.N = N
    for k in ("Vx", "Vu", "yref", "Vx_e", "yref_e", "W", "W_e", "zl", "Zl", "zu", "Zu"):
        setattr(ocp.cost, k, G["cost_" + k].copy())
    ocp.cost.zl_e = ocp.cost.zl.copy(); ocp.cost.Zl_e = ocp.cost.Zl.copy(); ocp.cost.zu_e = ocp.cost.zu.copy(); ocp.cost.Zu_e = ocp.cost.Zu.copy()
    for k in ("x0", "lbx", "ubx", "idxbx", "lbu", "ubu", "idxbu", "lh", "uh", "lh_e", "uh_e", "Jsh", "Jsh_e"):
        setattr(ocp.constraints, k, G["con_" + k].copy())
    ocp.parameter_values = np.zeros(10)
    o = ocp.solver_options
    o.qp_solver, o.hessian_approx, o.integrator_type, o.nlp_solver_type = "PARTIAL_CONDENSING_HPIPM", "GAUSS_NEWTON", "IRK", "SQP_RTI"
    o.levenberg_marquardt, o.qp_solver_iter_max, o.tf = 2.0, 50, tf
    return ocp


def test_ocp_translation_and_validation():
    ocp = _reference_ocp()
    cfg = acados_compat.ocp_to_config(ocp)
    assert cfg["N"] == 20 and cfg["n_obst"] == 5 and cfg["integrator"] == 0 and cfg["lm"] == 2.0 and cfg["qp_iter_max"] == 50
    assert cfg["W"] == [2, 2, 2, 2, 0.15, 0.15] and cfg["W_e"] == [5, 5, 5, 5]
    assert cfg["lbx"] == [-7, -7, -10, -10] and cfg["ubu"] == [8, 8] and cfg["uh"] == 1e5
    assert cfg["nlp_solver"] == 0                                           # 'SQP_RTI', robot_ocp_problem.py:130
    ocp.solver_options.nlp_solver_type = "SQP"                              # the full-SQP mode
    ocp.solver_options.nlp_solver_max_iter = 40; ocp.solver_options.nlp_solver_tol_stat = 1e-4
    ocp.solver_options.globalization = "MERIT_BACKTRACKING"
    cfg = acados_compat.ocp_to_config(ocp)
    assert cfg["nlp_solver"] == 1 and cfg["nlp_iter_max"] == 40 and cfg["nlp_tol"] == [1e-4, 1e-6, 1e-6, 1e-6] and cfg["nlp_globalization"] == 1
    ocp.solver_options.nlp_solver_type = "DDP"
    with pytest.raises(NotImplementedError):
        acados_compat.ocp_to_config(ocp)
    ocp = _reference_ocp(); ocp.solver_options.globalization = "FUNNEL_L1PEN_LINESEARCH"
    with pytest.raises(NotImplementedError):
        acados_compat.ocp_to_config(ocp)
    ocp = _reference_ocp(); ocp.constraints.idxbx = np.array([0, 1, 2, 3])
    with pytest.raises(NotImplementedError):
        acados_compat.ocp_to_config(ocp)
    ocp = _reference_ocp(); ocp.cost.cost_type = "NONLINEAR_LS"
    with pytest.raises(NotImplementedError):
        acados_compat.ocp_to_config(ocp)


def test_solver_facade_mirrors_and_errors():
    s = acados_compat.AcadosOcpSolver(_reference_ocp(), json_file="acados_ocp.json")
    s.set(3, "x", np.arange(5.0)); s.set(3, "u", [1.0, 2.0]); s.set(0, "ubx", np.ones(5)); s.set(0, "lbx", np.ones(5))
    s.cost_set(4, "zl", 7 * np.ones(5)); s.cost_set(4, "Zl", 8 * np.ones(5)); s.cost_set(20, "yref", [1, 2, 0, 0])
    s.set_params_sparse(2, np.arange(10), np.arange(10.0) + 1); s.set(2, "p", np.arange(10.0) + 1)
    assert np.array_equal(s.get(3, "x"), np.arange(5.0)) and np.array_equal(s.get(3, "u"), [1, 2])
    assert np.array_equal(s._f["x0"][0], np.ones(5)) and np.array_equal(s._f["zl"][0, 4], 7 * np.ones(5))
    assert np.array_equal(s._f["p"][0, 2], np.arange(10.0) + 1) and np.array_equal(s._f["yref"][0, 20, :4], [1, 2, 0, 0])
    got = s.get(3, "x"); got[:] = -1
    assert np.array_equal(s.get(3, "x"), np.arange(5.0))            # copy-out semantics
    with pytest.raises(Exception):
        s.set(3, "x", np.zeros(4))                                  # dimension mismatch
    with pytest.raises(Exception):
        s.cost_set(20, "yref", np.zeros(5))                         # the reference's set_subgoal bug (:284) raises
    with pytest.raises(Exception):
        s.set(20, "u", np.zeros(2))
    s.reset()
    assert not s.get(3, "x").any() and not s.get(3, "u").any()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "simulation")), reason="needs the reference checkout")
def test_reference_robot_ocp_problem_runs_unmodified_on_the_facade():
    """robot_ocp_problem.py (unmodified) constructs and parameterises its solver through the facade; what it
    wrote equals the golden vectors recorded from the same calls."""
    acados_compat.install()
    mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot"); mpl.animation = types.ModuleType("matplotlib.animation")
    sys.modules.setdefault("matplotlib", mpl); sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot); sys.modules.setdefault("matplotlib.animation", mpl.animation)
    for p in (os.path.join(REF, "src"), os.path.join(REF, "src", "simulation")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from robot_ocp_problem import RobotOcpProblem
    from models import world_specification as ws
    np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        prob = RobotOcpProblem(np.array([ws.X_MIN + 2, ws.Y_MIN + 2, np.pi / 4, 0.7, -0.3]), np.array([ws.X_MAX - 2, ws.Y_MAX - 2]),
                               scenario="RANDOM", slack=True, random_move=True)
    s = prob.ocp_solver
    assert isinstance(s, acados_compat.AcadosOcpSolver) and s.N == ws.N_SOLV and s.M == ws.N_OBST
    prob.x0 = G["slack_x0"].copy()
    prob.parameterize_slack()
    assert np.array_equal(s._f["zl"][0], G["slack_zl"]) and np.array_equal(s._f["Zl"][0], G["slack_Zl"])
    for o, row in zip(prob.obstacles, G["model_obst"]):
        o.x, o.y, o.vx, o.vy = row
    prob.parameterize_model()
    assert np.allclose(s._f["p"][0], G["model_p"], rtol=0, atol=1e-13)
    prob.x0 = np.array([1.0, -2.0, 0.3, 4.0, 5.0])
    prob.set_initial_guess()
    assert np.array_equal(s._f["x"][0], G["guess_x"]) and np.array_equal(prob.x0, G["guess_x0_after"])


@pytest.mark.gpu
def test_facade_closed_loop_equals_batched_solver(oracle):
    """drive the facade exactly like robot_ocp_problem.py:184-260 does (set p / cost_set zl,Zl / set lbx,ubx /
    solve / get u / sim / shift) and compare every step with the oracle on the same inputs"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    N, M, tf = 20, 5, 2.0
    dt = tf / N
    solver = acados_compat.AcadosOcpSolver(_reference_ocp(N, tf), json_file="acados_ocp.json")
    integrator = acados_compat.AcadosSimSolver(_reference_ocp(N, tf), json_file="acados_ocp.json")
    cfg = oracle.default_config(N, M, tf)
    obst, noise = world.reference_episode_streams(2, "RANDOM", M, 30, True)
    x0 = np.array([-6.0, -6.0, np.pi / 4, 0.0, 0.0]); goal = np.array([6.0, 6.0])
    yref = np.zeros((N + 1, 6)); yref[:, :2] = goal
    for i in range(N):
        solver.cost_set(i, "yref", yref[i])
    solver.cost_set(N, "yref", yref[N, :4])
    solver.reset()
    for i in range(N + 1):
        solver.set(i, "x", x0)
    xo = np.tile(x0, (N + 1, 1)); uo = np.zeros((N, 2)); slo = np.zeros((N + 1, M)); suo = np.zeros((N + 1, M))
    for step in range(12):
        P = world.predict_trajectory(obst, N, dt)
        zl, Zl = oracle.slack_weights(cfg, x0, goal)
        for i in range(N + 1):
            solver.set_params_sparse(i, np.arange(2 * M), P[i].flatten()); solver.set(i, "p", P[i].flatten())
            solver.cost_set(i, "zl", zl[i]); solver.cost_set(i, "Zl", Zl[i])
        solver.set(0, "ubx", x0); solver.set(0, "lbx", x0)
        stat = solver.solve()
        u = solver.get(0, "u")
        r = oracle.rti_solve(cfg, x0, np.ascontiguousarray(P.reshape(N + 1, 2 * M)), yref, zl, Zl, xo, uo, slo, suo)
        assert stat == r["status"] and abs(solver.get_stats("qp_iter") - r["qp_iter"]) <= 1
        tol = 1e-6 if (r["qp_iter"] < 50 and solver.get_stats("qp_iter") < 50) else 5e-2
        assert np.abs(u - uo[0]).max() < tol * max(1.0, np.abs(uo).max())
        assert np.abs(np.array([solver.get(i, "x") for i in range(N + 1)]) - xo).max() < tol * max(1.0, np.abs(xo).max())
        integrator.set("x", x0); integrator.set("u", uo[0]); integrator.solve()
        x0n = oracle.gl4_step(x0, uo[0].copy(), dt, sens=False)
        assert np.abs(integrator.get("x") - x0n).max() < 1e-12
        x0 = x0n
        obst = world.obstacle_step(obst, dt, True, noise[step])
        # shift exactly as the reference does (:253-258), through get/set; keep the facade on the oracle's iterate
        oracle.shift(cfg, xo, uo)
        for j in range(N + 1):
            solver.set(j, "x", xo[j]); solver.set(j, "sl", slo[j]); solver.set(j, "su", suo[j])
        for j in range(N):
            solver.set(j, "u", uo[j])


@pytest.mark.gpu
def test_facade_sqp_mode(oracle):
    """nlp_solver_type 'SQP' through the acados facade: statuses, get_stats('sqp_iter'), get_residuals(), reset()"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    N, M, tf = 20, 5, 2.0
    ocp = _reference_ocp(N, tf)
    o = ocp.solver_options
    o.nlp_solver_type, o.nlp_solver_max_iter = "SQP", 400
    o.nlp_solver_tol_stat = o.nlp_solver_tol_eq = o.nlp_solver_tol_ineq = o.nlp_solver_tol_comp = 1e-5
    solver = acados_compat.AcadosOcpSolver(ocp, json_file="acados_ocp.json")
    sc = world.random_scenarios(4, world.MpcDims(N, M, tf), seed=33)
    b = 1
    for i in range(N + 1):
        solver.set(i, "p", sc["p"][b, i]); solver.cost_set(i, "zl", sc["zl"][b, i]); solver.cost_set(i, "Zl", sc["Zl"][b, i])
        solver.cost_set(i, "yref", sc["yref"][b, i] if i < N else sc["yref"][b, i, :4])
        solver.set(i, "x", sc["x"][b, i])
    solver.set(0, "lbx", sc["x0"][b]); solver.set(0, "ubx", sc["x0"][b])
    stat = solver.solve()
    cfg = oracle.default_config(N, M, tf)
    w = {k: sc[k][b:b + 1].copy() for k in ("x", "u", "sl", "su")}
    r = oracle.sqp_solve_batch(cfg, sc["x0"][b:b + 1], sc["p"][b:b + 1], sc["yref"][b:b + 1], sc["zl"][b:b + 1], sc["Zl"][b:b + 1],
                               w["x"], w["u"], w["sl"], w["su"], max_iter=400, tol=(1e-5,) * 4)
    assert stat == r["status"][0] == 0
    assert abs(solver.get_stats("sqp_iter") - r["sqp_iter"][0]) <= 1 and solver.get_stats("sqp_iter") > 5
    assert (solver.get_residuals() < 1e-5).all()
    assert np.abs(solver.get(0, "u") - w["u"][0, 0]).max() < 1e-4
    assert solver.solve() == 0 and solver.get_stats("sqp_iter") == 0        # already converged: no QP
    keep = {k: solver._f[k].copy() for k in ("x", "u", "sl", "su")}
    solver.reset()                                                          # forgets the multipliers as well
    for k, v in keep.items():
        solver._f[k][:] = v
    assert solver.solve() == 0 and solver.get_stats("sqp_iter") >= 1        # stationarity must be re-established
