"""CPU tests of the MIRROR build (oracle/simt/): the CUDA kernel source compiled with g++ against a SIMT emulator.
The bit-for-bit comparison with the GPU is in tests/test_gpu_parity.py; here (no GPU) the mirror is checked against the
INDEPENDENT oracle (tolerances), which also exercises the kernel source itself in a container without a GPU."""
import ctypes as C

import numpy as np
import pytest

from b200mpc import capi, world


@pytest.fixture(scope="module")
def mirror():
    from oracle import pymirror
    pymirror.build()
    return pymirror


def _rel_rows(a, b):
    B = a.shape[0]
    return np.abs(a - b).reshape(B, -1).max(1) / np.maximum(1.0, np.abs(b).reshape(B, -1).max(1))


def test_mirror_defaults_equal_the_cabi_defaults(mirror):
    """pymirror.default_config spells the reference defaults out (it must work without libb200mpc.so's CUDA runtime):
    byte-identical to b2mpc_default_config"""
    a = mirror.default_config(20, 5, 2.0)
    b = capi.default_config(20, 5, 2.0, max_batch=1, device=0)
    assert bytes(a) == bytes(b)


def test_software_sincos_and_model_step(mirror, oracle):
    """csrc/b2_math.cuh (the sincos both the GPU and the mirror execute) inside the generated IRK-GL4 step, against the
    oracle's libm-based step"""
    rng = np.random.default_rng(0)
    for _ in range(200):
        x = np.array([rng.uniform(-7, 7), rng.uniform(-7, 7), rng.uniform(-40, 40), rng.uniform(-10, 10), rng.uniform(-10, 10)])
        u = rng.uniform(-8, 8, 2)
        xn, sv = mirror.gl4_step(x, u, 0.1)
        xo, A, Bm = oracle.gl4_step(x, u, 0.1)
        assert np.abs(xn - xo).max() < 1e-13 * max(1.0, np.abs(xo).max())
        ref = [A[0, 2], A[0, 3], A[0, 4], Bm[0, 0], Bm[0, 1], A[1, 2], A[1, 3], A[1, 4], Bm[1, 0], Bm[1, 1]]
        assert np.abs(sv - ref).max() < 1e-14


def test_mirror_layouts_are_bit_identical(mirror):
    """the persistent kernel (node-private rows in global memory) and the all-shared kernel differ in data placement only"""
    dims = world.MpcDims(20, 5, 2.0)
    sc = world.random_scenarios(24, dims, seed=3, moving_start=True)
    cfg = mirror.default_config(20, 5, 2.0)
    res = []
    for layout in (2, 3):
        m = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
        out = mirror.solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], m["x"], m["u"], m["sl"], m["su"], layout=layout)
        res.append((m, out))
    for k in ("x", "u", "sl", "su"):
        assert np.array_equal(res[0][0][k], res[1][0][k])
    for k in ("qp_iter", "status", "lam", "t", "pi"):
        assert np.array_equal(res[0][1][k], res[1][1][k])


@pytest.mark.parametrize("N,M,tf,B,qp_iter_max", [(20, 5, 2.0, 96, 10), (20, 5, 2.0, 96, 50), (10, 3, 1.0, 48, 50), (33, 2, 3.3, 16, 50)])
def test_mirror_matches_independent_oracle(mirror, oracle, N, M, tf, B, qp_iter_max):
    dims = world.MpcDims(N, M, tf)
    sc = world.random_scenarios(B, dims, seed=0)
    cfg = mirror.default_config(N, M, tf, qp_iter_max=qp_iter_max)
    m = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    out = mirror.solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], m["x"], m["u"], m["sl"], m["su"])
    ocfg = oracle.default_config(N, M, tf, qp_iter_max=qp_iter_max)
    r = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    d = {}
    st, it, stats = oracle.rti_solve_batch(ocfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], r["x"], r["u"], r["sl"], r["su"],
                                           nthreads=4, duals=d)
    assert (out["status"] == st).all()
    assert np.mean(out["qp_iter"] == it) >= 0.97 and np.abs(out["qp_iter"] - it).max() <= 1
    sel = (st == 0) & (np.ones(B, bool) if qp_iter_max < 50 else (it < 50) & (out["qp_iter"] < 50))
    assert sel.mean() > 0.5
    for k in ("u", "x"):
        assert _rel_rows(m[k], r[k])[sel].max() < 1e-6, k
    for k in ("lam", "t", "pi"):
        assert _rel_rows(out[k], d[k])[sel].max() < 1e-5, k


@pytest.mark.parametrize("N,tf,B", [(20, 2.0, 64), (10, 1.0, 48)])
def test_cta_kernel_mirror_matches_oracle_and_warp_kernel(mirror, oracle, N, tf, B):
    """the CTA-per-problem latency kernel (csrc/mpc_cta.cuh: 160-thread blocks, obstacle warps + node warp, feedback-form
    sweeps), executed by the emulator with block barriers and multi-warp reductions: the independent oracle's tolerance,
    and agreement with the warp-per-problem kernel to rounding"""
    dims = world.MpcDims(N, 5, tf)
    sc = world.random_scenarios(B, dims, seed=2, moving_start=True)
    cfg = mirror.default_config(N, 5, tf)
    res = {}
    for layout in (4, 2):
        m = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
        out = mirror.solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], m["x"], m["u"], m["sl"], m["su"], layout=layout)
        res[layout] = (m, out)
    r = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    d = {}
    st, it, _ = oracle.rti_solve_batch(oracle.default_config(N, 5, tf), sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"],
                                       r["x"], r["u"], r["sl"], r["su"], nthreads=4, duals=d)
    m4, o4 = res[4]
    assert (o4["status"] == st).all() and np.mean(o4["qp_iter"] == it) >= 0.95 and np.abs(o4["qp_iter"] - it).max() <= 1
    sel = (st == 0) & (it < 50) & (o4["qp_iter"] < 50)
    assert sel.mean() > 0.4
    for k in ("u", "x"):
        assert _rel_rows(m4[k], r[k])[sel].max() < 1e-6, k
        assert _rel_rows(m4[k], res[2][0][k])[sel].max() < 1e-6, k
    for k in ("lam", "pi"):
        assert _rel_rows(o4[k], d[k])[sel].max() < 1e-5, k
    assert np.mean(o4["qp_iter"] == res[2][1]["qp_iter"]) >= 0.95


def test_long_horizon_mirror_matches_oracle(mirror, oracle):
    """BASELINE config 4 (N = 100, 20 obstacles) on the CPU: the `<20,100,LONG>` instantiation -- node-private rows AND v, pi, dv,
    dpi in "global" memory, four node rounds per pass, the linear-residual pass fused into the expansion pass (FUSE_LIN), the
    software prefetch (a no-op in the emulator) -- against the independent oracle on a handful of problems."""
    N, M, tf, B = 100, 20, 5.0, 12
    dims = world.MpcDims(N, M, tf)
    sc = world.random_scenarios(B, dims, seed=8)
    m = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    out = mirror.solve_batch(mirror.default_config(N, M, tf), sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], m["x"], m["u"], m["sl"], m["su"])
    r = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    st, it, _ = oracle.rti_solve_batch(oracle.default_config(N, M, tf), sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"],
                                       r["x"], r["u"], r["sl"], r["su"], nthreads=4)
    assert (out["status"] == st).all()
    assert (out["qp_iter"] != it).sum() <= 1 and np.abs(out["qp_iter"] - it).max() <= 1
    sel = (st == 0) & (it < 50) & (out["qp_iter"] < 50)
    assert sel.sum() >= 2
    for k in ("u", "x"):
        assert _rel_rows(m[k], r[k])[sel].max() < 1e-6, k


@pytest.mark.parametrize("glob,min_same", [(0, 0.95), (1, 0.85)])
def test_sqp_mirror_matches_oracle(mirror, oracle, glob, min_same):
    """the full-SQP instantiation of the warp kernel (outer loop, convergence test, merit line search inside the kernel) on the
    CPU against the oracle's SQP loop: same statuses and SQP iteration counts up to the merit comparisons that rounding decides"""
    B, lm, max_iter = 24, 0.05, 15
    dims = world.MpcDims(20, 5, 2.0)
    sc = world.random_scenarios(B, dims, seed=14, moving_start=True)
    cfg = mirror.default_config(20, 5, 2.0, lm=lm, nlp_solver=1, nlp_iter_max=max_iter, nlp_globalization=glob)
    m = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    out = mirror.solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], m["x"], m["u"], m["sl"], m["su"], sqp=True)
    r = {k: sc[k].copy() for k in ("x", "u", "sl", "su")}
    ro = oracle.sqp_solve_batch(oracle.default_config(20, 5, 2.0, lm=lm), sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"],
                                r["x"], r["u"], r["sl"], r["su"], max_iter=max_iter, nthreads=4, globalization=glob)
    assert np.mean(out["status"] == ro["status"]) >= min_same
    assert np.mean(out["sqp_iter"] == ro["sqp_iter"]) >= min_same
    same = (out["status"] == ro["status"]) & (out["sqp_iter"] == ro["sqp_iter"])
    e = _rel_rows(m["u"], r["u"])
    assert np.median(e[same]) < 1e-8
