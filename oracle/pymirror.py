"""ctypes binding of the CPU MIRROR of the CUDA solve kernels (oracle/simt/wpp_mirror.cpp).

TEST INFRASTRUCTURE ONLY.  The mirror is the kernel SOURCE of dynamic-obstacle-avoidance-mpc_b200/csrc/ compiled with g++
against a small SIMT emulator (oracle/simt/simt_emu.h): same operations, same order, same reductions as the GPU, so tests
demand bit-identical results (`np.array_equal`), not tolerances.  It says nothing about whether the algorithm is right --
that is what the independent oracle (oracle/mpc_oracle.c, `pyoracle`) is for.  Only tests/, __graft_entry__.smoke() and
bench.py's checker legs import this module; the CUDA product never does.

Arrays are numpy float64, batch-first, like pyoracle's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_CSRC = os.path.join(_ROOT, "dynamic-obstacle-avoidance-mpc_b200", "csrc")
_LIB_PATH = os.path.join(_HERE, "_build", "libwpp_mirror.so")
_SRC = os.path.join(_HERE, "simt", "wpp_mirror.cpp")
CXXFLAGS = ["-O2", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", "-march=x86-64-v3", "-fPIC", "-shared",
            "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-unused-but-set-variable"]


def _deps():
    d = [_SRC, os.path.join(_HERE, "simt", "simt_emu.h"), os.path.join(_ROOT, "include", "b200mpc.h")]
    d += [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cuh", ".h"))]
    return d


def build(force: bool = False) -> str:
    """g++ build of the kernel source against the emulator (recipe also in oracle/Makefile: `make mirror`)."""
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in _deps())
    if force or stale:
        subprocess.check_call(["g++"] + CXXFLAGS + ["-I", os.path.join(_HERE, "simt"), "-o", _LIB_PATH, _SRC, "-lpthread"])
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _config_type():
    pkg = os.path.join(_ROOT, "dynamic-obstacle-avoidance-mpc_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from b200mpc import capi          # the ctypes struct of include/b200mpc.h: b2mpc_config (no GPU needed to import)
    return capi.Config


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.b2mirror_solve.restype = C.c_int
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _pi(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


# reference defaults, spelled out here so that the mirror needs neither a GPU nor libb200mpc.so
# (tests/test_cabi_cpu.py checks them against b2mpc_default_config)
def default_config(N=20, n_obst=5, tf=2.0, **overrides):
    Config = _config_type()
    c = Config()
    c.N = N; c.n_obst = n_obst; c.device = 0; c.max_batch = 1
    c.integrator = 0; c.qp_iter_max = 50; c.lm_scaled_by_dt = 1; c.kernel = 0
    c.pred_corr = 1; c.cond_pred_corr = 1; c.itref_corr_max = 2
    c.tf = tf; c.lm = 2.0
    for i in range(4):
        c.W[i] = 2.0; c.W_e[i] = 5.0
    c.W[4] = c.W[5] = 0.15
    c.lbx[0] = c.lbx[1] = -7.0; c.lbx[2] = c.lbx[3] = -10.0
    c.ubx[0] = c.ubx[1] = 7.0; c.ubx[2] = c.ubx[3] = 10.0
    c.lbu[0] = c.lbu[1] = -8.0; c.ubu[0] = c.ubu[1] = 8.0
    c.r2 = 5.7599999999999998; c.lh = 0.0; c.uh = 1e5; c.zu = 0.0; c.Zu = 0.0
    c.mu0 = 1.0; c.alpha_min = 1e-8
    c.res_g_max = 1e-6; c.res_b_max = 1e-8; c.res_d_max = 1e-8; c.res_m_max = 1e-8
    c.reg_prim = 1e-15; c.lam_min = 1e-16; c.t_min = 1e-16; c.tau_min = 1e-16; c.thr0 = 0.1
    c.nlp_solver = 0; c.nlp_iter_max = 100
    for i in range(4):
        c.nlp_tol[i] = 1e-6
    c.nlp_globalization = 0; c.ls_alpha_min = 0.05; c.ls_alpha_reduction = 0.7; c.ls_eps = 0.0
    for k, v in overrides.items():
        cur = getattr(c, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(c, k, v)
    return c


def solve_batch(cfg, x0, p, yref, zl, Zl, x, u, sl, su, layout=2, sqp=False, pi=None, lam=None, t=None, nthreads=None):
    """One b2mpc_solve() of the batch through the mirrored kernel; iterate (x, u, sl, su) updated IN PLACE.
    layout 2 = the persistent throughput kernel (node-private rows in global memory), 3 = all-shared kernel.
    Returns dict(status, qp_iter, qp_status, stats, pi, lam, t[, sqp_iter, nlp_res])."""
    B = x0.shape[0]; N, M = cfg.N, cfg.n_obst
    nct = 12 + 4 * M
    pi = np.zeros((B, N, 5)) if pi is None else pi
    lam = np.zeros((B, N + 1, nct)) if lam is None else lam
    t = np.zeros((B, N + 1, nct)) if t is None else t
    stats = np.zeros((B, 8)); nlp_res = np.zeros((B, 4))
    status = np.zeros(B, np.int32); qp_iter = np.zeros(B, np.int32); qp_status = np.zeros(B, np.int32); sqp_iter = np.zeros(B, np.int32)
    if nthreads is None:
        nthreads = min(os.cpu_count() or 1, 32)
    rc = lib().b2mirror_solve(C.byref(cfg), C.c_int(layout), C.c_int(1 if sqp else 0), C.c_int64(B), C.c_int(nthreads),
                              _p(x0), _p(p), _p(yref), _p(zl), _p(Zl), _p(x), _p(u), _p(sl), _p(su),
                              _p(pi), _p(lam), _p(t), _p(stats), _p(nlp_res),
                              _pi(status), _pi(qp_iter), _pi(qp_status), _pi(sqp_iter))
    if rc != 0:
        raise ValueError("b2mirror_solve: unsupported configuration")
    out = dict(status=status, qp_iter=qp_iter, qp_status=qp_status, stats=stats, pi=pi, lam=lam, t=t)
    if sqp:
        out.update(sqp_iter=sqp_iter, nlp_res=nlp_res)
    return out


def gl4_step(x, u, h):
    x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
    xn = np.zeros(5); sv = np.zeros(10)
    lib().b2mirror_gl4_step(_p(x), _p(u), C.c_double(h), _p(xn), _p(sv))
    return xn, sv
