"""Generic implicit Runge-Kutta (Gauss-Legendre collocation) integrator with Newton
iterations and implicit-function-theorem forward sensitivities -- the algorithm acados'
`sim_irk_integrator` runs for `integrator_type = 'IRK'` (robot_ocp_problem.py:129) with
its defaults (4 stages, 1 step, 3 Newton iterations from K = 0).

ORACLE / TEST INFRASTRUCTURE ONLY (parity unpinned, see oracle/mpc_oracle.h).  It exists
to PROVE that the closed form in oracle/mpc_oracle.c:orc_gl4_step (and the generated CUDA
device function) equals what the reference's integrator computes for the model of
src/models/robot_model.py:39-43.
"""
import numpy as np


def gauss_legendre_tableau(s: int):
    """Butcher tableau (A, b, c) of the s-stage Gauss-Legendre collocation method on [0,1]."""
    z, w = np.polynomial.legendre.leggauss(s)
    c = 0.5 * (z + 1.0)
    b = 0.5 * w
    # a_ij = integral_0^{c_i} l_j(tau) dtau, l_j = Lagrange basis on the nodes c
    A = np.zeros((s, s))
    for j in range(s):
        others = [c[m] for m in range(s) if m != j]
        poly = np.poly1d([1.0])
        for cm in others:
            poly = poly * np.poly1d([1.0, -cm]) / (c[j] - cm)
        P = poly.integ()
        for i in range(s):
            A[i, j] = P(c[i]) - P(0.0)
    return A, b, c


def f_expl(x, u):
    """robot_model.py:39-43"""
    return np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), x[4], u[0], u[1]])


def f_jac(x, u):
    fx = np.zeros((5, 5)); fu = np.zeros((5, 2))
    fx[0, 2] = -x[3] * np.sin(x[2]); fx[0, 3] = np.cos(x[2])
    fx[1, 2] = x[3] * np.cos(x[2]);  fx[1, 3] = np.sin(x[2])
    fx[2, 4] = 1.0
    fu[3, 0] = 1.0; fu[4, 1] = 1.0
    return fx, fu


def irk_step(x, u, h, stages=4, newton_iter=3):
    """One IRK step: returns x_next, dx_next/dx, dx_next/du."""
    A, b, c = gauss_legendre_tableau(stages)
    nx, nu, s = 5, 2, stages
    K = np.zeros((s, nx))
    for _ in range(newton_iter):
        G = np.zeros(s * nx)
        J = np.eye(s * nx)
        for i in range(s):
            xi = x + h * A[i] @ K
            G[i * nx:(i + 1) * nx] = K[i] - f_expl(xi, u)
            fx, _ = f_jac(xi, u)
            for j in range(s):
                J[i * nx:(i + 1) * nx, j * nx:(j + 1) * nx] -= h * A[i, j] * fx
        K = K - np.linalg.solve(J, G).reshape(s, nx)
    xn = x + h * b @ K
    # IFT sensitivities at the final K:  J dK/dw = [fx_i, fu_i]
    J = np.eye(s * nx)
    R = np.zeros((s * nx, nx + nu))
    for i in range(s):
        xi = x + h * A[i] @ K
        fx, fu = f_jac(xi, u)
        R[i * nx:(i + 1) * nx, :nx] = fx
        R[i * nx:(i + 1) * nx, nx:] = fu
        for j in range(s):
            J[i * nx:(i + 1) * nx, j * nx:(j + 1) * nx] -= h * A[i, j] * fx
    dK = np.linalg.solve(J, R).reshape(s, nx, nx + nu)
    S = h * np.einsum("i,ijk->jk", b, dK)
    Ax = np.eye(nx) + S[:, :nx]
    Bu = S[:, nx:]
    return xn, Ax, Bu
