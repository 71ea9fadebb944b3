"""CPU checks of the oracle's full-SQP mode (north_star kernel 4: outer loop with convergence test; the
reference itself runs SQP_RTI, robot_ocp_problem.py:130).  The NLP residuals are pinned by an independent
finite-difference gradient of the Lagrangian written in numpy; the loop by its fixed-point properties."""
import numpy as np

from b200mpc import world

DIMS = world.MpcDims(20, 5, 2.0)


def _scen(B, seed=3, moving=True):
    return world.random_scenarios(B, DIMS, seed=seed, moving_start=moving)


def _lagrangian(oracle, cfg, sc, b, x, u, sl, su, pi, lam, t):
    """cost + pi.(F - x+) - lam.g  with g_i >= 0 the inequality functions, padded layout [lb6 lg ub6 ug ls us]."""
    N, M = cfg.N, cfg.M
    dt = cfg.tf / N
    W = np.array(list(cfg.W)); We = np.array(list(cfg.We))
    idx = [0, 1, 3, 4]
    L = 0.0
    for k in range(N + 1):
        yr = sc["yref"][b, k]
        scal = dt if k < N else 1.0
        if k < N:
            y = np.concatenate([x[k, idx], u[k]])
            L += 0.5 * scal * np.sum(W * (y - yr) ** 2)
            xn = oracle.gl4_step(x[k], u[k], dt, sens=False)
            L += pi[k] @ (xn - x[k + 1])
            for s in range(2):
                L -= lam[k, s] * (u[k, s] - cfg.lbu[s]) + lam[k, 6 + M + s] * (cfg.ubu[s] - u[k, s])
            if k > 0:
                for s in range(4):
                    L -= lam[k, 2 + s] * (x[k, idx[s]] - cfg.lbx[s]) + lam[k, 8 + M + s] * (cfg.ubx[s] - x[k, idx[s]])
        else:
            L += 0.5 * np.sum(We * (x[k, idx] - yr[:4]) ** 2)
        h, _, _ = oracle.con_h(x[k], sc["p"][b, k], cfg.r2)
        zl, Zl = sc["zl"][b, k], sc["Zl"][b, k]
        L += scal * np.sum(zl * sl[k] + 0.5 * Zl * sl[k] ** 2)
        L -= lam[k, 6:6 + M] @ (h + sl[k] - cfg.lh) + lam[k, 12 + M:12 + 2 * M] @ (cfg.uh - h + su[k])
        L -= lam[k, 12 + 2 * M:12 + 3 * M] @ sl[k] + lam[k, 12 + 3 * M:12 + 4 * M] @ su[k]
    return L


def test_nlp_residuals_match_a_finite_difference_lagrangian(oracle):
    sc = _scen(2)
    cfg = oracle.default_config(DIMS.N, DIMS.M, DIMS.tf)
    N, M = DIMS.N, DIMS.M
    for b in range(2):
        x, u, sl, su = (sc[k][b].copy() for k in ("x", "u", "sl", "su"))
        for _ in range(3):                       # a few SQP steps: non-trivial multipliers and slacks
            o = oracle.rti_solve(cfg, sc["x0"][b], sc["p"][b], sc["yref"][b], sc["zl"][b], sc["Zl"][b], x, u, sl, su, want_duals=True)
        pi, lam, t = o["pi"], o["lam"], o["t"]
        res = oracle.nlp_residuals(cfg, sc["x0"][b], sc["p"][b], sc["yref"][b], sc["zl"][b], sc["Zl"][b], x, u, sl, su, pi, lam, t)
        # stationarity: central differences over u, x_1.., sl, su
        g = []
        f = lambda: _lagrangian(oracle, cfg, sc, b, x, u, sl, su, pi, lam, t)
        for arr, ks in ((u, range(N)), (x, range(1, N + 1)), (sl, range(N + 1)), (su, range(N + 1))):
            for k in ks:
                for i in range(arr.shape[1]):
                    v0 = arr[k, i]; hstep = 1e-5 * max(1.0, abs(v0))
                    arr[k, i] = v0 + hstep; fp = f()
                    arr[k, i] = v0 - hstep; fm = f()
                    arr[k, i] = v0
                    g.append((fp - fm) / (2 * hstep))
        fd = np.max(np.abs(g))
        scale = max(1.0, np.abs(lam).max())
        assert abs(fd - res[0]) <= 1e-6 * scale + 1e-4 * res[0], (fd, res[0])
        # dynamics gap, inequality residual and complementarity, recomputed in numpy
        gap = max(np.abs(oracle.gl4_step(x[k], u[k], DIMS.dt, sens=False) - x[k + 1]).max() for k in range(N))
        assert abs(gap - res[1]) < 1e-14
        assert abs(np.abs(lam * t).max() - res[3]) < 1e-12 * max(1.0, res[3])
        h = np.stack([oracle.con_h(x[k], sc["p"][b, k], cfg.r2)[0] for k in range(N + 1)])
        ineq = max(np.abs(h + sl - cfg.lh - t[:, 6:6 + M]).max(), np.abs(cfg.uh - h + su - t[:, 12 + M:12 + 2 * M]).max(),
                   np.abs(sl - t[:, 12 + 2 * M:12 + 3 * M]).max(), np.abs(su - t[:, 12 + 3 * M:]).max(),
                   np.abs(u - np.array(list(cfg.lbu)) - t[:N, 0:2]).max(), np.abs(np.array(list(cfg.ubu)) - u - t[:N, 6 + M:8 + M]).max(),
                   np.abs(x[1:N, [0, 1, 3, 4]] - np.array(list(cfg.lbx)) - t[1:N, 2:6]).max(),
                   np.abs(np.array(list(cfg.ubx)) - x[1:N, [0, 1, 3, 4]] - t[1:N, 8 + M:12 + M]).max(),
                   np.abs(sc["x0"][b] - x[0]).max())
        assert abs(ineq - res[2]) < 1e-9 * max(1.0, ineq)


def test_sqp_with_one_iteration_is_the_rti_step(oracle):
    B = 16
    sc = _scen(B)
    cfg = oracle.default_config(DIMS.N, DIMS.M, DIMS.tf)
    a = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    st, it, _ = oracle.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], a["x"], a["u"], a["sl"], a["su"])
    c = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    r = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], c["x"], c["u"], c["sl"], c["su"], max_iter=1)
    for k in a:
        assert np.array_equal(a[k], c[k])                      # bit-identical: the same QP, the same full step
    assert np.array_equal(r["qp_iter"], it)
    assert set(np.unique(r["status"])) <= {2, 4} and np.array_equal(r["status"] == 4, st == 4)   # ACADOS_MAXITER unless the QP failed
    assert (r["sqp_iter"] == 1).all()


def test_sqp_converges_and_is_idempotent(oracle):
    B = 48
    sc = _scen(B, seed=5)
    cfg = oracle.default_config(DIMS.N, DIMS.M, DIMS.tf)
    it = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    r = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], it["x"], it["u"], it["sl"], it["su"],
                               max_iter=400, nthreads=8)
    ok = r["status"] == 0
    assert ok.mean() > 0.6                                     # LM = 2 makes the Gauss-Newton iteration linear (rate ~0.9)
    assert (r["res"][ok] < 1e-6).all()
    assert (r["sqp_iter"][ok] < 400).all() and (r["sqp_iter"][r["status"] == 2] == 400).all()
    # converged problems: x_0 = x0bar, dynamics closed, a second call does nothing (bit-identical iterate, 0 QPs)
    assert np.abs(it["x"][ok, 0] - sc["x0"][ok]).max() == 0.0
    before = {k: v.copy() for k, v in it.items()}
    r2 = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], it["x"], it["u"], it["sl"], it["su"],
                                pi=r["pi"], lam=r["lam"], t=r["t"], max_iter=5, nthreads=8)
    assert (r2["sqp_iter"][ok] == 0).all() and (r2["status"][ok] == 0).all()
    for k in it:
        assert np.array_equal(it[k][ok], before[k][ok])
    # ... and WITHOUT the multipliers (reset) the first test fails on stationarity but one QP restores them
    it3 = {k: before[k].copy() for k in before}
    r3 = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], it3["x"], it3["u"], it3["sl"], it3["su"],
                                max_iter=5, nthreads=8)
    assert (r3["sqp_iter"][ok] >= 1).all()
    assert np.abs(it3["u"][ok] - before["u"][ok]).max() < 1e-4


def test_looser_tolerance_needs_fewer_iterations(oracle):
    B = 16
    sc = _scen(B, seed=7)
    cfg = oracle.default_config(DIMS.N, DIMS.M, DIMS.tf)
    its = []
    for tol in (1e-2, 1e-4):
        w = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
        r = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], w["x"], w["u"], w["sl"], w["su"],
                                   max_iter=300, tol=(tol,) * 4, nthreads=8)
        its.append(r["sqp_iter"])
    assert (its[0] <= its[1]).all() and its[0].mean() < its[1].mean()


def test_merit_backtracking_helps_where_full_steps_oscillate(oracle):
    """globalization = MERIT_BACKTRACKING: with the reference's lm = 2 the Levenberg-Marquardt term already damps the
    steps (the line search changes little); with a small lm full Gauss-Newton steps oscillate and the line search
    rescues a good part of the problems."""
    B = 96
    sc = _scen(B, seed=1)
    conv = {}
    for lm, mi in ((2.0, 300), (0.05, 100)):
        cfg = oracle.default_config(DIMS.N, DIMS.M, DIMS.tf, lm=lm)
        for g in (0, 1):
            w = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
            r = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], w["x"], w["u"], w["sl"], w["su"],
                                       max_iter=mi, nthreads=8, globalization=g)
            conv[(lm, g)] = np.mean(r["status"] == 0)
            assert (r["res"][r["status"] == 0] < 1e-6).all()
    assert abs(conv[(2.0, 1)] - conv[(2.0, 0)]) <= 0.05
    assert conv[(0.05, 1)] >= conv[(0.05, 0)] + 0.1, conv


def test_full_line_search_steps_equal_fixed_steps(oracle):
    """one SQP iteration from the cold iterate with an accepted alpha = 1 is bit-identical to the FIXED_STEP iteration"""
    B = 32
    sc = _scen(B, seed=2)
    cfg = oracle.default_config(DIMS.N, DIMS.M, DIMS.tf)
    a = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    ra = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], a["x"], a["u"], a["sl"], a["su"], max_iter=1)
    c = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    rc = oracle.sqp_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], c["x"], c["u"], c["sl"], c["su"], max_iter=1, globalization=1)
    full = np.array([np.array_equal(a["u"][b], c["u"][b]) for b in range(B)])
    assert full.mean() > 0.5                                  # most first steps are accepted in full ...
    for b in np.nonzero(full)[0]:
        assert np.array_equal(a["x"][b], c["x"][b]) and np.array_equal(ra["lam"][b], rc["lam"][b])
    for b in np.nonzero(~full)[0]:                            # ... the others moved along the same direction, but less far
        if ra["status"][b] == 4:
            continue
        d_full = a["u"][b] - sc["u"][b]; d_ls = c["u"][b] - sc["u"][b]
        al = (d_full * d_ls).sum() / (d_full * d_full).sum()
        assert 0.049 <= al < 1.0 and np.abs(d_ls - al * d_full).max() < 1e-9 * max(1.0, np.abs(d_full).max())
