"""ctypes binding of the CPU ORACLE (oracle/mpc_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/mpc_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import
this module.  The CUDA product (dynamic-obstacle-avoidance-mpc_b200/) never does.

All arrays are numpy float64, problem-major / stage-major ("batch-first"):
  x[B,N+1,5] u[B,N,2] sl,su[B,N+1,M] p[B,N+1,2M] yref[B,N+1,6] zl,Zl[B,N+1,M] x0[B,5]
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmpc_oracle.so")
_LIB_PATH_V3 = os.path.join(_HERE, "_build", "libmpc_oracle_v3.so")


def _cpu_has_v3() -> bool:
    """x86-64-v3 build (AVX2/FMA/BMI2 code generation, FP contraction still off) only where the CPU has it"""
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
        return all(f in flags for f in ("avx2", "fma", "bmi2", "movbe"))
    except Exception:
        return False


class OrcConfig(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("M", C.c_int32), ("integrator", C.c_int32), ("soft", C.c_int32),
        ("qp_iter_max", C.c_int32), ("lm_scaled_by_dt", C.c_int32),
        ("pred_corr", C.c_int32), ("cond_pred_corr", C.c_int32),
        ("itref_corr_max", C.c_int32), ("lq_fact", C.c_int32),
        ("tf", C.c_double), ("lm", C.c_double),
        ("W", C.c_double * 6), ("We", C.c_double * 4),
        ("lbx", C.c_double * 4), ("ubx", C.c_double * 4),
        ("lbu", C.c_double * 2), ("ubu", C.c_double * 2),
        ("r2", C.c_double), ("lh", C.c_double), ("uh", C.c_double),
        ("zu", C.c_double), ("Zu", C.c_double),
        ("mu0", C.c_double), ("alpha_min", C.c_double),
        ("res_g_max", C.c_double), ("res_b_max", C.c_double),
        ("res_d_max", C.c_double), ("res_m_max", C.c_double),
        ("reg_prim", C.c_double), ("lam_min", C.c_double), ("t_min", C.c_double),
        ("tau_min", C.c_double), ("thr0", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/mpc_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "mpc_oracle.c")
    hdr = os.path.join(_HERE, "mpc_oracle.h")
    stale = any((not os.path.exists(lp)) or any(os.path.getmtime(f) > os.path.getmtime(lp) for f in (src, hdr))
                for lp in (_LIB_PATH, _LIB_PATH_V3))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH_V3 if _cpu_has_v3() else _LIB_PATH)
        _lib.orc_rti_solve.restype = C.c_int
        _lib.orc_export_qp.restype = C.c_int
        _lib.orc_qp_solve_export.restype = C.c_int
        _lib.orc_sqp_solve.restype = C.c_int
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _pi(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def default_config(N=20, M=5, tf=2.0, **kw) -> OrcConfig:
    c = OrcConfig()
    lib().orc_default_config(C.byref(c), C.c_int(N), C.c_int(M), C.c_double(tf))
    for k, v in kw.items():
        cur = getattr(c, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(c, k, v)
    return c


# ------------------------------------------------------------------ model
def gl4_step(x, u, h, sens=True):
    x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
    xn = np.zeros(5); A = np.zeros((5, 5)); B = np.zeros((5, 2))
    lib().orc_gl4_step(_p(x), _p(u), C.c_double(h), _p(xn), _p(A) if sens else None, _p(B) if sens else None)
    return (xn, A, B) if sens else xn


def rk4_step(x, u, h, sens=True):
    x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
    xn = np.zeros(5); A = np.zeros((5, 5)); B = np.zeros((5, 2))
    lib().orc_rk4_step(_p(x), _p(u), C.c_double(h), _p(xn), _p(A) if sens else None, _p(B) if sens else None)
    return (xn, A, B) if sens else xn


def con_h(x, p, r2):
    x = np.ascontiguousarray(x, np.float64); p = np.ascontiguousarray(p, np.float64)
    M = p.size // 2
    h = np.zeros(M); cx = np.zeros(M); cy = np.zeros(M)
    lib().orc_con_h(_p(x), _p(p), C.c_int(M), C.c_double(r2), _p(h), _p(cx), _p(cy))
    return h, cx, cy


# ------------------------------------------------------------------ solve
def rti_solve(cfg, x0, p, yref, zl, Zl, x, u, sl, su, want_duals=False):
    """One problem, arrays without batch dim; iterate updated IN PLACE."""
    N, M = cfg.N, cfg.M
    it = C.c_int32(0); qs = C.c_int32(0)
    stats = np.zeros(8)
    pi = np.zeros((N, 5)) if want_duals else None
    lam = np.zeros((N + 1, 12 + 4 * M)) if want_duals else None
    t = np.zeros((N + 1, 12 + 4 * M)) if want_duals else None
    st = lib().orc_rti_solve(C.byref(cfg), _p(x0), _p(p), _p(yref), _p(zl), _p(Zl),
                             _p(x), _p(u), _p(sl), _p(su), _p(pi), _p(lam), _p(t),
                             C.byref(it), C.byref(qs), _p(stats))
    out = dict(status=st, qp_iter=it.value, qp_status=qs.value, stats=stats)
    if want_duals:
        out.update(pi=pi, lam=lam, t=t)
    return out


def rti_solve_batch(cfg, x0, p, yref, zl, Zl, x, u, sl, su, nthreads=1, duals=None):
    """Batch-first arrays; iterate updated IN PLACE; returns (status, qp_iter, stats).  `duals`: an optional dict that
    receives the multipliers pi[B,N,5], lam[B,N+1,12+4M], t[B,N+1,12+4M] of the QP."""
    B = x0.shape[0]
    status = np.zeros(B, np.int32); qp_iter = np.zeros(B, np.int32); stats = np.zeros((B, 8))
    if duals is not None:
        N, M = cfg.N, cfg.M
        duals.update(pi=np.zeros((B, N, 5)), lam=np.zeros((B, N + 1, 12 + 4 * M)), t=np.zeros((B, N + 1, 12 + 4 * M)))
        lib().orc_rti_solve_batch_duals(C.byref(cfg), C.c_int64(B), C.c_int(nthreads),
                                        _p(x0), _p(p), _p(yref), _p(zl), _p(Zl), _p(x), _p(u), _p(sl), _p(su),
                                        _p(duals["pi"]), _p(duals["lam"]), _p(duals["t"]), _pi(status), _pi(qp_iter), _p(stats))
        return status, qp_iter, stats
    lib().orc_rti_solve_batch(C.byref(cfg), C.c_int64(B), C.c_int(nthreads),
                              _p(x0), _p(p), _p(yref), _p(zl), _p(Zl),
                              _p(x), _p(u), _p(sl), _p(su), _pi(status), _pi(qp_iter), _p(stats))
    return status, qp_iter, stats


# ---------------------------------------------------- full SQP (north_star kernel 4)
NLP_TOL = (1e-6, 1e-6, 1e-6, 1e-6)          # acados nlp_solver_tol_{stat,eq,ineq,comp} defaults


def nlp_residuals(cfg, x0, p, yref, zl, Zl, x, u, sl, su, pi, lam, t):
    """[stationarity, dynamics gap, inequality, complementarity] inf norms of one problem."""
    res = np.zeros(4)
    lib().orc_nlp_residuals(C.byref(cfg), _p(x0), _p(p), _p(yref), _p(zl), _p(Zl), _p(x), _p(u), _p(sl), _p(su),
                            _p(pi), _p(lam), _p(t), _p(res))
    return res


def sqp_solve_batch(cfg, x0, p, yref, zl, Zl, x, u, sl, su, pi=None, lam=None, t=None, max_iter=100, tol=NLP_TOL, nthreads=1, globalization=0,
                    ls=(0.05, 0.7, 0.0)):
    """Batch-first arrays; iterate AND multipliers updated in place.  Returns dict(status, sqp_iter, qp_iter, res, pi, lam, t,
    qp_maxed = number of QPs per problem that ran into qp_iter_max: such QPs are not at a KKT point, what follows is path dependent)."""
    B = x0.shape[0]; N, M = cfg.N, cfg.M
    pi = np.zeros((B, N, 5)) if pi is None else pi
    lam = np.zeros((B, N + 1, 12 + 4 * M)) if lam is None else lam
    t = np.zeros((B, N + 1, 12 + 4 * M)) if t is None else t
    status = np.zeros(B, np.int32); sqp_iter = np.zeros(B, np.int32); qp_iter = np.zeros(B, np.int32); res = np.zeros((B, 4))
    tol_a = np.ascontiguousarray(tol, np.float64); ls_a = np.ascontiguousarray(ls, np.float64)
    qp_maxed = np.zeros(B, np.int32)
    lib().orc_sqp_solve_batch(C.byref(cfg), C.c_int64(B), C.c_int(nthreads), C.c_int(max_iter), _p(tol_a), C.c_int(globalization), _p(ls_a),
                              _p(x0), _p(p), _p(yref), _p(zl), _p(Zl), _p(x), _p(u), _p(sl), _p(su), _p(pi), _p(lam), _p(t),
                              _pi(status), _pi(sqp_iter), _pi(qp_iter), _p(res), _pi(qp_maxed))
    return dict(status=status, sqp_iter=sqp_iter, qp_iter=qp_iter, res=res, pi=pi, lam=lam, t=t, qp_maxed=qp_maxed)


def export_qp(cfg, x0, p, yref, zl, Zl, x, u, sl, su):
    """Dense RTI QP of one problem (what the IPM sees; node 0 has x0 eliminated).
    v-order per node [u0,u1,x,y,psi,v,w]; BAt[k,i,r] = d x+_r / d v_i."""
    N, M = cfg.N, cfg.M
    q = dict(BAt=np.zeros((N + 1, 7, 5)), b=np.zeros((N + 1, 5)), H=np.zeros((N + 1, 7, 7)),
             g=np.zeros((N + 1, 7)), Cxy=np.zeros((N + 1, M, 2)),
             lb=np.zeros((N + 1, 6)), ub=np.zeros((N + 1, 6)),
             lg=np.zeros((N + 1, M)), ug=np.zeros((N + 1, M)),
             ls=np.zeros((N + 1, M)), us=np.zeros((N + 1, M)),
             Zl=np.zeros((N + 1, M)), Zu=np.zeros((N + 1, M)),
             zl=np.zeros((N + 1, M)), zu=np.zeros((N + 1, M)))
    lib().orc_export_qp(C.byref(cfg), _p(x0), _p(p), _p(yref), _p(zl), _p(Zl), _p(x), _p(u), _p(sl), _p(su),
                        _p(q["BAt"]), _p(q["b"]), _p(q["H"]), _p(q["g"]), _p(q["Cxy"]),
                        _p(q["lb"]), _p(q["ub"]), _p(q["lg"]), _p(q["ug"]), _p(q["ls"]), _p(q["us"]),
                        _p(q["Zl"]), _p(q["Zu"]), _p(q["zl"]), _p(q["zu"]))
    return q


def qp_solve_export(cfg, x0, p, yref, zl, Zl, x, u, sl, su):
    N, M = cfg.N, cfg.M
    s = dict(v=np.zeros((N + 1, 7)), dsl=np.zeros((N + 1, M)), dsu=np.zeros((N + 1, M)),
             pi=np.zeros((N, 5)), lam=np.zeros((N + 1, 12 + 4 * M)), t=np.zeros((N + 1, 12 + 4 * M)))
    it = C.c_int32(0); qs = C.c_int32(0); stats = np.zeros(8)
    lib().orc_qp_solve_export(C.byref(cfg), _p(x0), _p(p), _p(yref), _p(zl), _p(Zl), _p(x), _p(u), _p(sl), _p(su),
                              _p(s["v"]), _p(s["dsl"]), _p(s["dsu"]), _p(s["pi"]), _p(s["lam"]), _p(s["t"]),
                              C.byref(it), C.byref(qs), _p(stats))
    s.update(qp_iter=it.value, qp_status=qs.value, stats=stats)
    return s


# ------------------------------------------------------------------- glue
def slack_weights(cfg, x0, goal):
    N, M = cfg.N, cfg.M
    zl = np.zeros((N + 1, M)); Zl = np.zeros((N + 1, M))
    lib().orc_slack_weights(C.byref(cfg), _p(np.ascontiguousarray(x0, np.float64)),
                            _p(np.ascontiguousarray(goal, np.float64)), _p(zl), _p(Zl))
    return zl, Zl


def shift(cfg, x, u):
    lib().orc_shift(C.byref(cfg), _p(x), _p(u))


def initial_guess(cfg, x0, x, u, sl, su):
    lib().orc_initial_guess(C.byref(cfg), _p(x0), _p(x), _p(u), _p(sl), _p(su))


X_MIN, X_MAX, Y_MIN, Y_MAX = -8.0, 8.0, -8.0, 8.0      # world_specification.py:7-10


def obst_predict_traj(o, n, dt, quirk=True):
    o = np.ascontiguousarray(o, np.float64)
    traj = np.zeros((n + 1, 2))
    lib().orc_obst_predict_traj(_p(o), C.c_int(n), C.c_double(dt), C.c_int(1 if quirk else 0),
                                C.c_double(X_MIN), C.c_double(X_MAX), C.c_double(Y_MIN), C.c_double(Y_MAX), _p(traj))
    return traj


def obst_step(o, dt, random_move, noise, randomness=0.1, vmax=2.0):
    """in place on o = [x,y,vx,vy]"""
    noise = np.ascontiguousarray(noise, np.float64)
    lib().orc_obst_step(_p(o), C.c_double(dt), C.c_int(1 if random_move else 0), C.c_double(randomness),
                        C.c_double(vmax), _p(noise),
                        C.c_double(X_MIN), C.c_double(X_MAX), C.c_double(Y_MIN), C.c_double(Y_MAX))
