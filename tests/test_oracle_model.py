"""Oracle discretisation checks (CPU): the closed-form GL4 step of oracle/mpc_oracle.c equals
the generic IRK-Newton scheme acados runs, a tight-tolerance ODE solve,
and finite differences; one-step RK4 does NOT meet the 1e-6 parity bar."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from oracle import irk_generic


def _samples(n, seed=0, vmax=10.0, wmax=10.0):
    rng = np.random.default_rng(seed)
    x = np.stack([rng.uniform(-7, 7, n), rng.uniform(-7, 7, n), rng.uniform(-np.pi, np.pi, n),
                  rng.uniform(-vmax, vmax, n), rng.uniform(-wmax, wmax, n)], axis=1)
    u = rng.uniform(-8, 8, (n, 2))
    return x, u


def test_tableau_is_gauss_legendre_4():
    A, b, c = irk_generic.gauss_legendre_tableau(4)
    assert np.allclose(c, [0.06943184420297371, 0.33000947820757187, 0.6699905217924281, 0.9305681557970262], atol=1e-15)
    assert np.allclose(b, [0.17392742256872679, 0.3260725774312732, 0.3260725774312732, 0.17392742256872679], atol=1e-15)
    assert np.allclose(A.sum(1), c, atol=1e-14)
    assert np.allclose(A @ c, c ** 2 / 2, atol=1e-14)          # C(2), used by the closed form


@pytest.mark.parametrize("h", [0.1, 0.05])
def test_gl4_closed_form_equals_irk_newton3(oracle, h):
    x, u = _samples(300, seed=1)
    err_x = err_s = 0.0
    for xi, ui in zip(x, u):
        xn, A, B = oracle.gl4_step(xi, ui, h)
        xr, Ar, Br = irk_generic.irk_step(xi, ui, h, stages=4, newton_iter=3)
        err_x = max(err_x, np.abs(xn - xr).max())
        err_s = max(err_s, np.abs(A - Ar).max(), np.abs(B - Br).max())
    assert err_x < 2e-14, err_x
    assert err_s < 5e-15, err_s


def test_irk_newton_converged_after_two_iterations(oracle):
    x, u = _samples(50, seed=2)
    for xi, ui in zip(x, u):
        x2, _, _ = irk_generic.irk_step(xi, ui, 0.1, newton_iter=2)
        x3, _, _ = irk_generic.irk_step(xi, ui, 0.1, newton_iter=3)
        assert np.abs(x2 - x3).max() < 1e-14


def test_gl4_matches_tight_ode_solve(oracle):
    x, u = _samples(40, seed=3, vmax=5.0, wmax=2.0)
    for xi, ui in zip(x, u):
        xn = oracle.gl4_step(xi, ui, 0.1, sens=False)
        sol = solve_ivp(lambda t, s: irk_generic.f_expl(s, ui), (0, 0.1), xi, rtol=1e-13, atol=1e-14)
        assert np.abs(xn - sol.y[:, -1]).max() < 1e-9     # GL4 is order 8


def test_gl4_sensitivities_vs_central_differences(oracle):
    x, u = _samples(40, seed=4)
    eps = 1e-6
    for xi, ui in zip(x, u):
        _, A, B = oracle.gl4_step(xi, ui, 0.1)
        for j in range(5):
            d = np.zeros(5); d[j] = eps
            fd = (oracle.gl4_step(xi + d, ui, 0.1, sens=False) - oracle.gl4_step(xi - d, ui, 0.1, sens=False)) / (2 * eps)
            assert np.abs(A[:, j] - fd).max() < 1e-7
        for j in range(2):
            d = np.zeros(2); d[j] = eps
            fd = (oracle.gl4_step(xi, ui + d, 0.1, sens=False) - oracle.gl4_step(xi, ui - d, 0.1, sens=False)) / (2 * eps)
            assert np.abs(B[:, j] - fd).max() < 1e-7


def test_rk4_sensitivities_and_gap_to_irk(oracle):
    """RK4 (north_star's named integrator) is self-consistent but misses the reference's IRK
    by far more than the 1e-6 parity bar -> never parity-test RK4 against IRK numbers."""
    x, u = _samples(200, seed=5)
    gap = 0.0
    eps = 1e-6
    for xi, ui in zip(x, u):
        xr, A, B = oracle.rk4_step(xi, ui, 0.1)
        gap = max(gap, np.abs(xr - oracle.gl4_step(xi, ui, 0.1, sens=False)).max())
    assert 1e-5 < gap < 1e-2, gap
    xi, ui = x[0], u[0]
    _, A, B = oracle.rk4_step(xi, ui, 0.1)
    for j in range(5):
        d = np.zeros(5); d[j] = eps
        fd = (oracle.rk4_step(xi + d, ui, 0.1, sens=False) - oracle.rk4_step(xi - d, ui, 0.1, sens=False)) / (2 * eps)
        assert np.abs(A[:, j] - fd).max() < 1e-7
    for j in range(2):
        d = np.zeros(2); d[j] = eps
        fd = (oracle.rk4_step(xi, ui + d, 0.1, sens=False) - oracle.rk4_step(xi, ui - d, 0.1, sens=False)) / (2 * eps)
        assert np.abs(B[:, j] - fd).max() < 1e-7


def test_constraint_and_gradient(oracle):
    rng = np.random.default_rng(6)
    x = rng.uniform(-7, 7, 5); p = rng.uniform(-6, 6, 10)
    h, cx, cy = oracle.con_h(x, p, 5.76)
    for j in range(5):
        assert h[j] == pytest.approx((x[0] - p[2 * j]) ** 2 + (x[1] - p[2 * j + 1]) ** 2 - 2.4 ** 2, abs=1e-13)
        assert cx[j] == pytest.approx(2 * (x[0] - p[2 * j])) and cy[j] == pytest.approx(2 * (x[1] - p[2 * j + 1]))
