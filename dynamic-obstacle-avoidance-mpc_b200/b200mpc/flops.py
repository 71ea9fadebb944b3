"""Algorithmic FLOP / byte model of one RTI solve (frozen).  Dense counts on
purpose -- what HPIPM/BLASFEO would execute; FMA = 2 flops.  Used by bench.py for the roofline
numerator of every kernel variant, so fractions stay comparable across variants."""

NX, NU = 5, 2
T_SINCOS = 40


def counts(N: int, M: int):
    nx, nu = NX, NU
    nv = nx + nu
    nb = 6
    ng = ns = M
    n_c = 2 * nb + 2 * ng + 2 * ns
    f_dyn = 4 * (T_SINCOS + 30) + 40
    f_node = 9 * M + 3 * nv
    f_lin = N * f_dyn + (N + 1) * f_node
    forward = 2 * nu * nx + nu * nu + 2 * nx * nv + 2 * nx * nx + nx
    per_node = (12 * (nb + ng) + 16 * ns
                + 2 * nb + ng * nv * (nv + 1) + 3 * ng * nv
                + (nv + 1) * nx * nx + nx + nx * nv * (nv + 3) + nv ** 3 / 3 + nv * nv
                + forward
                + 2 * (nx * nx + nx * nv + nv * nv) + forward
                + 2 * (2 * ng * nv + 10 * n_c) + 8 * n_c + 6 * n_c + 2 * (nv + nx + 2 * n_c)
                + 2 * nv * nv + 4 * nx * nv + 4 * ng * nv + 8 * n_c)
    f_ipm = (N + 1) * per_node
    nbytes = 8 * (nx + 2 * M * (N + 1) + 2 + 2 * ((N + 1) * nx + N * nu)) + 8
    return dict(f_lin=f_lin, f_ipm=f_ipm, bytes=nbytes, per_node=per_node)


def flops_per_solve(N: int, M: int, mean_qp_iter: float) -> float:
    c = counts(N, M)
    return c["f_lin"] + mean_qp_iter * c["f_ipm"]
