"""Oracle QP checks (CPU).  The RTI QP is convex, so a point that satisfies the KKT conditions
to tolerance IS its solution: we rebuild the KKT residuals in numpy from the exported dense QP
(independent of the oracle's Riccati/elimination code) and check the oracle's IPM output."""
import numpy as np
import pytest

from b200mpc import world


def _kkt(q, s, N, M):
    """max-norm KKT residuals of solution s on exported QP q (node 0 has x0 eliminated)."""
    v, pi, lam, t = s["v"], s["pi"], s["lam"], s["t"]
    dsl, dsu = s["dsl"], s["dsu"]
    o_lb, o_lg, o_ub, o_ug, o_ls, o_us = 0, 6, 6 + M, 12 + M, 12 + 2 * M, 12 + 3 * M
    slot2v = [0, 1, 2, 3, 5, 6]          # box slots [u0,u1,x,y,v,w] -> index in [u;x]
    r_stat = r_dyn = r_pf = r_comp = r_sl = 0.0
    for k in range(N + 1):
        vk = v[k].copy()
        if k == 0:
            vk[2:] = 0.0                 # eliminated: its effect is folded into b_0 / lg_0 / ug_0
        grad = q["H"][k] @ vk + q["g"][k]
        if k < N:
            grad += q["BAt"][k] @ pi[k]
        if k > 0:
            grad[2:] -= pi[k - 1]
        C = np.zeros((M, 7)); C[:, 2:4] = q["Cxy"][k]
        grad -= C.T @ (lam[k, o_lg:o_lg + M] - lam[k, o_ug:o_ug + M])
        for sl_, iv in enumerate(slot2v):
            if np.isfinite(q["lb"][k, sl_]):
                grad[iv] += -lam[k, o_lb + sl_] + lam[k, o_ub + sl_]
                r_pf = max(r_pf, abs(vk[iv] - q["lb"][k, sl_] - t[k, o_lb + sl_]), abs(q["ub"][k, sl_] - vk[iv] - t[k, o_ub + sl_]))
        rows = slice(0, 2) if k == 0 else (slice(2, 7) if k == N else slice(0, 7))
        r_stat = max(r_stat, np.abs(grad[rows]).max())
        cv = C @ vk
        r_pf = max(r_pf, np.abs(cv + dsl[k] - q["lg"][k] - t[k, o_lg:o_lg + M]).max(),
                   np.abs(q["ug"][k] - cv + dsu[k] - t[k, o_ug:o_ug + M]).max(),
                   np.abs(dsl[k] - q["ls"][k] - t[k, o_ls:o_ls + M]).max(),
                   np.abs(dsu[k] - q["us"][k] - t[k, o_us:o_us + M]).max())
        r_sl = max(r_sl, np.abs(q["Zl"][k] * dsl[k] + q["zl"][k] - lam[k, o_lg:o_lg + M] - lam[k, o_ls:o_ls + M]).max(),
                   np.abs(q["Zu"][k] * dsu[k] + q["zu"][k] - lam[k, o_ug:o_ug + M] - lam[k, o_us:o_us + M]).max())
        r_comp = max(r_comp, np.abs(lam[k] * t[k]).max())
        assert (lam[k] >= 0).all() and (t[k] >= 0).all()
        if k < N:
            xn = q["BAt"][k].T @ vk + q["b"][k]
            r_dyn = max(r_dyn, np.abs(xn - v[k + 1][2:]).max())
    return r_stat, r_sl, r_dyn, r_pf, r_comp


def _closed_loop_states(oracle, cfg, sc, b, steps):
    """advance problem b of scenario set sc a few closed-loop steps so constraints get active"""
    N, M = cfg.N, cfg.M
    dt = cfg.tf / N
    x0 = sc["x0"][b].copy(); x = sc["x"][b].copy(); u = sc["u"][b].copy()
    sl = sc["sl"][b].copy(); su = sc["su"][b].copy()
    obst = sc["obst"][b].copy(); goal = sc["goal"][b]
    for _ in range(steps):
        p = np.ascontiguousarray(world.predict_trajectory(obst, N, dt).reshape(N + 1, 2 * M))
        zl, Zl = oracle.slack_weights(cfg, x0, goal)
        oracle.rti_solve(cfg, x0, p, sc["yref"][b], zl, Zl, x, u, sl, su)
        x0 = oracle.gl4_step(x0, u[0].copy(), dt, sens=False)
        obst = world.obstacle_step(obst, dt)
        oracle.shift(cfg, x, u)
    p = np.ascontiguousarray(world.predict_trajectory(obst, N, dt).reshape(N + 1, 2 * M))
    zl, Zl = oracle.slack_weights(cfg, x0, goal)
    return x0, p, zl, Zl, x, u, sl, su


@pytest.mark.parametrize("N,M,tf", [(20, 5, 2.0), (5, 5, 0.5), (10, 3, 1.0)])
def test_ipm_solution_satisfies_kkt(oracle, N, M, tf):
    cfg = oracle.default_config(N, M, tf)
    sc = world.random_scenarios(12, world.MpcDims(N, M, tf), seed=3, moving_start=True)
    n_conv = 0
    for b in range(12):
        args = _closed_loop_states(oracle, cfg, sc, b, steps=(b % 4) * 6)
        x0, p, zl, Zl, x, u, sl, su = args
        q = oracle.export_qp(cfg, x0, p, sc["yref"][b], zl, Zl, x, u, sl, su)
        s = oracle.qp_solve_export(cfg, x0, p, sc["yref"][b], zl, Zl, x, u, sl, su)
        assert s["qp_status"] in (0, 1)
        if s["qp_status"] != 0:
            continue
        n_conv += 1
        r_stat, r_sl, r_dyn, r_pf, r_comp = _kkt(q, s, N, M)
        assert r_stat < 2e-6 and r_sl < 2e-6, (r_stat, r_sl)     # res_g_max = 1e-6
        assert r_dyn < 1e-7 and r_pf < 1e-7, (r_dyn, r_pf)        # res_b/d_max = 1e-8 (+ rounding)
        assert r_comp < 1.1e-8, r_comp                            # res_m_max = 1e-8
    assert n_conv >= 9


def test_qp_data_matches_survey_appendix_a(oracle):
    """Hessian/gradient diagonals of SURVEY Appendix A: H_k = dt*diag(.15,.15,2,2,0,2,2)+2I."""
    cfg = oracle.default_config(20, 5, 2.0, lm_scaled_by_dt=0)      # SURVEY App. A quotes the unscaled variant
    sc = world.random_scenarios(1, world.MpcDims(), seed=0)
    a = [sc[k][0] for k in ("x0", "p", "yref", "zl", "Zl", "x", "u", "sl", "su")]
    q = oracle.export_qp(cfg, *a)
    assert np.allclose(np.diag(q["H"][3]), [2.015, 2.015, 2.2, 2.2, 2.0, 2.2, 2.2])
    assert np.allclose(np.diag(q["H"][20])[2:], [7, 7, 2, 7, 7])
    assert np.allclose(np.diag(q["H"][0])[:2], [2.015, 2.015]) and np.allclose(q["H"][0][2:, 2:], 0)
    x = sc["x"][0]
    assert np.allclose(q["g"][3][2:4], 0.1 * 2 * (x[3, :2] - 6.0))
    assert np.allclose(q["g"][20][2:4], 5 * (x[20, :2] - 6.0))
    assert np.allclose(q["lb"][3], [-8, -8, -7 - x[3, 0], -7 - x[3, 1], -10, -10])
    assert np.isinf(q["lb"][20]).all() and np.isinf(q["lb"][0][2:]).all()
    zl = sc["zl"][0]
    assert np.allclose(q["Zl"][3], 0.1 * zl[3]) and np.allclose(q["zl"][3], 0.1 * zl[3])
    assert np.allclose(q["Zl"][20], 0.0)
    cfg2 = oracle.default_config(20, 5, 2.0)                         # default: Ts * levenberg_marquardt
    assert cfg2.lm_scaled_by_dt == 1
    q2 = oracle.export_qp(cfg2, *a)
    assert np.allclose(np.diag(q2["H"][3]), [0.215, 0.215, 0.4, 0.4, 0.2, 0.4, 0.4])


def test_rti_step_updates_iterate_and_batch_equals_single(oracle):
    cfg = oracle.default_config(20, 5, 2.0)
    sc = world.random_scenarios(16, world.MpcDims(), seed=11, moving_start=True)
    xs, us, sls, sus = (sc[k].copy() for k in ("x", "u", "sl", "su"))
    status, qp_iter, stats = oracle.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"],
                                                    xs, us, sls, sus, nthreads=4)
    assert (status == 0).all()
    assert qp_iter.min() >= 5 and qp_iter.max() <= 50
    for b in range(16):
        x, u, sl, su = (sc[k][b].copy() for k in ("x", "u", "sl", "su"))
        r = oracle.rti_solve(cfg, sc["x0"][b], sc["p"][b], sc["yref"][b], sc["zl"][b], sc["Zl"][b], x, u, sl, su)
        assert r["qp_iter"] == qp_iter[b]
        assert np.array_equal(x, xs[b]) and np.array_equal(u, us[b])
        assert np.allclose(x[0], sc["x0"][b], atol=1e-15)          # x0 embedding
        assert (np.abs(u) <= 8 + 1e-7).all()
        assert (np.abs(x[1:20, [0, 1]]) <= 7 + 1e-7).all()


def test_glue_shift_and_initial_guess(oracle):
    cfg = oracle.default_config(5, 5, 0.5)
    x = np.arange(30, dtype=float).reshape(6, 5); u = np.arange(10, dtype=float).reshape(5, 2) + 100
    x_ref, u_ref = x.copy(), u.copy()
    oracle.shift(cfg, x, u)
    assert np.array_equal(x[:4], x_ref[1:5]) and np.array_equal(x[4], x_ref[5]) and np.array_equal(x[5], x_ref[5])
    assert np.array_equal(u[:4], u_ref[1:5]) and np.array_equal(u[4], [0, 0])
    x0 = np.array([1.0, 2.0, 0.3, 4.0, 5.0]); sl = np.ones((6, 5)); su = np.ones((6, 5))
    oracle.initial_guess(cfg, x0, x, u, sl, su)
    assert np.array_equal(x0, [1, 2, 0.3, 0, 0])                   # aliasing quirk, :301-302
    assert (x == x0).all() and (u == 0).all() and (sl == 0).all() and (su == 0).all()


def test_lq_fact_variants_agree(oracle):
    """HPIPM's lq_fact switch restated in the oracle (round 2): 0 = Cholesky of the summed node matrix, 1 = switch
    to the LQ-based factorisation when the predictor's linear residual exceeds 1e-5 (mode BALANCE), 2 = LQ always.  The LQ
    factor of [chol(H) | sqrt(Gamma_b) | C sqrt(Gamma_g) | AL] is the Cholesky factor of the same node matrix, so the three
    variants walk the same interior-point path up to rounding: same solutions on converged QPs, iteration counts equal but
    for the +-1 of the exit test."""
    from b200mpc import world
    dims = world.MpcDims(20, 5, 2.0)
    B = 96
    sc = world.random_scenarios(B, dims, seed=7, moving_start=True)
    res = {}
    for lq in (0, 1, 2):
        cfg = oracle.default_config(20, 5, 2.0, lq_fact=lq)
        it_ = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
        st, it, stats = oracle.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], it_["x"], it_["u"], it_["sl"], it_["su"], nthreads=4)
        res[lq] = (it_, st, it, stats)
    assert (res[0][3][:, 7] == 0).all()                      # stats[7]: iterations that ran on the LQ factorisation
    assert (res[2][3][:, 7] == res[2][2]).all()              # lq_fact = 2: all of them
    sw = res[1][3][:, 7] > 0
    assert 0.1 < sw.mean() < 0.9                             # mode BALANCE switches on a good part of the cold problems, not on all
    for lq in (1, 2):
        assert (res[lq][1] == res[0][1]).all()
        assert np.mean(res[lq][2] == res[0][2]) >= 0.95 and np.abs(res[lq][2] - res[0][2]).max() <= 2
        conv = (res[lq][2] < 50) & (res[0][2] < 50) & (res[0][1] == 0)
        du = np.abs(res[lq][0]["u"] - res[0][0]["u"]).reshape(B, -1).max(1)
        assert conv.mean() > 0.4 and du[conv].max() < 1e-6
