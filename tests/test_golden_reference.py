"""Golden vectors produced by EXECUTING the reference's own Python code (tests/golden/make_golden.py):
obstacle model, scenario sampling, OCP definition, per-step parameterisation, initial guess, shift and the
model expressions.  Checked here (CPU) against the host-side mirror (b200mpc.world), the oracle and the
generated model code; tests/test_gpu_glue.py checks the CUDA kernels against the same vectors."""
import ctypes as C
import os

import numpy as np
import pytest

from b200mpc import world

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
DT = 0.1


def test_world_constants():
    w = G["world_constants"]
    mine = [world.X_MIN, world.X_MAX, world.Y_MIN, world.Y_MAX, world.R_ROBOT, world.V_MAX_ROBOT, world.C_MAX, world.N_OBST,
            world.R_OBST, world.RANDOMNESS, world.V_MAX_OBST, world.MARGIN, world.X_MIN_OBST, world.X_MAX_OBST, world.Y_MIN_OBST,
            world.Y_MAX_OBST, world.TOL, world.QP_ITER, world.X_MIN_ROBOT, world.X_MAX_ROBOT]
    assert np.allclose(w, mine, rtol=0, atol=1e-15)


@pytest.mark.parametrize("scen", ["RANDOM", "EDGE", "CENTER"])
def test_scenario_sampling_and_noise_stream_are_reference_identical(scen):
    init, traj = G[f"obst_init_{scen}"], G[f"obst_traj_{scen}"]
    for seed in range(init.shape[0]):
        obst, noise = world.reference_episode_streams(seed, scen, 5, traj.shape[1], True)
        assert np.array_equal(obst, init[seed])
        o = obst.copy()
        for k in range(traj.shape[1]):
            o = world.obstacle_step(o, DT, True, noise[k])
            assert np.allclose(o, traj[seed, k], rtol=0, atol=1e-13), (seed, k)


@pytest.mark.parametrize("scen", ["RANDOM", "EDGE"])
def test_horizon_prediction_incl_vx_quirk(scen, oracle):
    init, pred0 = G[f"obst_init_{scen}"], G[f"obst_pred0_{scen}"]
    for seed in range(init.shape[0]):
        P = world.predict_trajectory(init[seed], 20, DT, quirk=True)         # [21, M, 2]
        assert np.allclose(np.transpose(P, (1, 0, 2)), pred0[seed], rtol=0, atol=1e-13)
        for j in range(5):
            assert np.allclose(oracle.obst_predict_traj(init[seed, j], 20, DT, quirk=True), pred0[seed, j], rtol=0, atol=1e-13)
        # without the quirk the prediction differs (the reference really predicts with vx = vy)
        assert not np.allclose(world.predict_trajectory(init[seed], 20, DT, quirk=False)[..., 0], P[..., 0])
    # mid-episode prediction
    traj, predm = G[f"obst_traj_{scen}"], G[f"obst_predmid_{scen}"]
    for seed in range(init.shape[0]):
        P = world.predict_trajectory(traj[seed, traj.shape[1] // 2], 20, DT)
        assert np.allclose(np.transpose(P, (1, 0, 2)), predm[seed], rtol=0, atol=1e-13)


def test_oracle_obstacle_step_matches_reference(oracle):
    init, traj = G["obst_init_RANDOM"], G["obst_traj_RANDOM"]
    for seed in range(3):
        _, noise = world.reference_episode_streams(seed, "RANDOM", 5, traj.shape[1], True)
        o = init[seed].copy()
        for k in range(traj.shape[1]):
            for j in range(5):
                oracle.obst_step(o[j], DT, True, noise[k, j])
            assert np.allclose(o, traj[seed, k], rtol=0, atol=1e-13)


def test_ocp_definition_matches_default_config(oracle):
    """robot_ocp_problem.py:54-132 as built by the reference vs the constants baked into the configs"""
    cfg = oracle.default_config(20, 5, 2.0)
    assert np.allclose(np.diag(G["cost_W"]), list(cfg.W)) and np.count_nonzero(G["cost_W"] - np.diag(np.diag(G["cost_W"]))) == 0
    assert np.allclose(np.diag(G["cost_W_e"]), list(cfg.We))
    # y = Vx x + Vu u selects [x, y, v, omega, u_a, u_alpha]
    Vx, Vu = G["cost_Vx"], G["cost_Vu"]
    sel = np.zeros((6, 5)); sel[0, 0] = sel[1, 1] = sel[2, 3] = sel[3, 4] = 1
    assert np.array_equal(Vx, sel) and np.array_equal(Vu[4:], np.eye(2)) and not Vu[:4].any()
    selx = np.zeros((4, 5)); selx[0, 0] = selx[1, 1] = selx[2, 3] = selx[3, 4] = 1
    assert np.array_equal(G["cost_Vx_e"], selx)
    assert np.allclose(G["con_lbx"], list(cfg.lbx)) and np.allclose(G["con_ubx"], list(cfg.ubx))
    assert np.array_equal(G["con_idxbx"], [0, 1, 3, 4]) and np.array_equal(G["con_idxbu"], [0, 1])
    assert np.allclose(G["con_lbu"], list(cfg.lbu)) and np.allclose(G["con_ubu"], list(cfg.ubu))
    assert np.allclose(G["con_lh"], cfg.lh) and np.allclose(G["con_uh"], cfg.uh)
    assert np.allclose(G["con_lh_e"], cfg.lh) and np.allclose(G["con_uh_e"], cfg.uh)
    assert np.array_equal(G["con_Jsh"], np.eye(5)) and np.array_equal(G["con_Jsh_e"], np.eye(5))
    assert np.allclose(G["cost_zu"], cfg.zu) and np.allclose(G["cost_Zu"], cfg.Zu)
    lm, tf, qpit = G["opt_lm_tf_qpiter"]
    assert lm == cfg.lm and qpit == cfg.qp_iter_max
    assert list(G["opt_strings"]) == ["PARTIAL_CONDENSING_HPIPM", "GAUSS_NEWTON", "IRK", "SQP_RTI", "LINEAR_LS", "LINEAR_LS"]
    assert cfg.integrator == 0          # IRK is the parity integrator


def test_parameterize_slack_matches_reference(oracle):
    N = int(G["ocp_N"][0])
    cfg = oracle.default_config(N, 5, 0.1 * N)
    zl, Zl = oracle.slack_weights(cfg, G["slack_x0"], G["slack_goal"])
    assert np.allclose(zl, G["slack_zl"], rtol=1e-15, atol=0) and np.allclose(Zl, G["slack_Zl"], rtol=1e-15, atol=0)
    zl2, Zl2 = world.slack_weights(G["slack_x0"][None], G["slack_goal"][None], N, 5)
    assert np.allclose(zl2[0], G["slack_zl"], rtol=1e-15, atol=0) and np.allclose(Zl2[0], G["slack_Zl"], rtol=1e-15, atol=0)


def test_parameterize_model_layout_matches_reference():
    N = int(G["ocp_N"][0])
    P = world.predict_trajectory(G["model_obst"], N, DT, quirk=True)         # [N+1, M, 2]
    assert np.allclose(P.reshape(N + 1, 10), G["model_p"], rtol=0, atol=1e-13)


def test_initial_guess_and_shift_match_reference(oracle):
    N = int(G["ocp_N"][0])
    cfg = oracle.default_config(N, 5, 0.1 * N)
    x0 = np.array([1.0, -2.0, 0.3, 4.0, 5.0])
    x = np.zeros((N + 1, 5)); u = np.ones((N, 2)); sl = np.ones((N + 1, 5)); su = np.ones((N + 1, 5))
    oracle.initial_guess(cfg, x0, x, u, sl, su)
    assert np.array_equal(x0, G["guess_x0_after"])            # aliasing quirk (:301-302)
    assert np.array_equal(x, G["guess_x"]) and np.array_equal(u, G["guess_u"])
    X, U = G["shift_X_in"].copy(), G["shift_U_in"].copy()
    oracle.shift(cfg, X, U)
    assert np.array_equal(X, G["shift_X_out"]) and np.array_equal(U, G["shift_U_out"])


def test_model_expressions_oracle_and_generated_code(oracle):
    from b200mpc import build_ext
    lib = C.CDLL(build_ext.build_model_host())
    dp = C.POINTER(C.c_double)
    lib.b2model_gl4_pattern.restype = C.c_char_p
    lib.b2model_r2.restype = C.c_double
    assert lib.b2model_gl4_pattern().decode() == "10vvvvv/01vvvvv/0010c0c/00010c0/000010c"
    assert abs(lib.b2model_r2() - (world.R_OBST + world.R_ROBOT + world.MARGIN) ** 2) < 1e-15
    import ctypes
    for x, u, pp, f_ref, h_ref in zip(G["model_x"], G["model_u"], G["model_pp"], G["model_f"], G["model_h"]):
        # oracle ODE right-hand side (robot_model.py:39-43) and constraints (:60-65)
        f = np.zeros(5)
        oracle.lib().orc_ode(x.ctypes.data_as(dp), u.ctypes.data_as(dp), f.ctypes.data_as(dp))
        assert np.allclose(f, f_ref, rtol=0, atol=1e-14)
        h, cx, cy = oracle.con_h(x, pp, 5.76)
        assert np.allclose(h, h_ref, rtol=0, atol=1e-12)
        # generated device/host code: constraint + integrator against the oracle's closed form
        for j in range(5):
            hv, gx, gy = C.c_double(), C.c_double(), C.c_double()
            lib.b2model_con_h(x.ctypes.data_as(dp), C.c_double(pp[2 * j]), C.c_double(pp[2 * j + 1]), C.c_double(5.76),
                              C.byref(hv), C.byref(gx), C.byref(gy))
            assert abs(hv.value - h_ref[j]) < 1e-12 and abs(gx.value - cx[j]) < 1e-13 and abs(gy.value - cy[j]) < 1e-13
        for step, ostep in ((lib.b2model_gl4_step, oracle.gl4_step), (lib.b2model_rk4_step, oracle.rk4_step)):
            xn = np.zeros(5); sv = np.zeros(10)
            step(x.ctypes.data_as(dp), u.ctypes.data_as(dp), C.c_double(DT), xn.ctypes.data_as(dp), sv.ctypes.data_as(dp))
            xo, A, B = ostep(x, u, DT)
            assert np.allclose(xn, xo, rtol=0, atol=5e-14)
            ref_sv = [A[0, 2], A[0, 3], A[0, 4], B[0, 0], B[0, 1], A[1, 2], A[1, 3], A[1, 4], B[1, 0], B[1, 1]]
            assert np.allclose(sv, ref_sv, rtol=0, atol=5e-14)
            # the constant part of the pattern
            assert A[2, 4] == DT and B[3, 0] == DT and B[4, 1] == DT and abs(B[2, 1] - 0.5 * DT * DT) < 1e-18
