"""ctypes binding of libb200mpc.so (include/b200mpc.h).  The library is built in-tree by
`build_ext.build()` (nvcc, sm_100a) and there is NO fallback: a missing library or a missing
CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libb200mpc.so")

# field ids (include/b200mpc.h)
X0, P, YREF, ZL, ZZL, X, U, SL, SU, PI, LAM, T, STATS, GOAL, OBST, MARGIN, U0, NLP_RES = range(18)
STATUS, QP_ITER, QP_STATUS, STEPS, DONE, OOB, NFAIL, SQP_ITER = range(8)
LAYOUT_SOA, LAYOUT_AOS = 0, 1
FIELD_NAMES = {"x0": X0, "p": P, "yref": YREF, "zl": ZL, "Zl": ZZL, "x": X, "u": U, "sl": SL, "su": SU,
               "pi": PI, "lam": LAM, "t": T, "stats": STATS, "goal": GOAL, "obst": OBST, "margin": MARGIN, "u0": U0,
               "nlp_res": NLP_RES}
IFIELD_NAMES = {"status": STATUS, "qp_iter": QP_ITER, "qp_status": QP_STATUS, "steps": STEPS, "done": DONE, "oob": OOB, "nfail": NFAIL,
                "sqp_iter": SQP_ITER}


class Config(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("n_obst", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32),
        ("integrator", C.c_int32), ("qp_iter_max", C.c_int32), ("lm_scaled_by_dt", C.c_int32), ("kernel", C.c_int32),
        ("pred_corr", C.c_int32), ("cond_pred_corr", C.c_int32), ("itref_corr_max", C.c_int32), ("reserved", C.c_int32),
        ("tf", C.c_double), ("lm", C.c_double),
        ("W", C.c_double * 6), ("W_e", C.c_double * 4),
        ("lbx", C.c_double * 4), ("ubx", C.c_double * 4), ("lbu", C.c_double * 2), ("ubu", C.c_double * 2),
        ("r2", C.c_double), ("lh", C.c_double), ("uh", C.c_double), ("zu", C.c_double), ("Zu", C.c_double),
        ("mu0", C.c_double), ("alpha_min", C.c_double), ("res_g_max", C.c_double), ("res_b_max", C.c_double),
        ("res_d_max", C.c_double), ("res_m_max", C.c_double),
        ("reg_prim", C.c_double), ("lam_min", C.c_double), ("t_min", C.c_double), ("tau_min", C.c_double), ("thr0", C.c_double),
        ("nlp_solver", C.c_int32), ("nlp_iter_max", C.c_int32), ("nlp_tol", C.c_double * 4),
        ("nlp_globalization", C.c_int32), ("reserved2", C.c_int32), ("ls_alpha_min", C.c_double), ("ls_alpha_reduction", C.c_double),
        ("ls_eps", C.c_double),
    ]


class B2mpcError(RuntimeError):
    pass


_lib = None

EXPORTS = [
    "b2mpc_abi_version", "b2mpc_last_error", "b2mpc_default_config", "b2mpc_create", "b2mpc_destroy",
    "b2mpc_field_len", "b2mpc_workspace_bytes", "b2mpc_set", "b2mpc_get", "b2mpc_get_i32",
    "b2mpc_set_host", "b2mpc_get_host", "b2mpc_get_i32_host", "b2mpc_field_ptr",
    "b2mpc_prepare_slack", "b2mpc_reset", "b2mpc_solve", "b2mpc_sim", "b2mpc_shift",
    "b2mpc_launch_count", "b2mpc_kernel", "b2mpc_solve_host", "b2mpc_dfma_peak",
    "b2mpc_obst_predict", "b2mpc_loop_reset", "b2mpc_loop_post", "b2mpc_solve_sqp", "b2mpc_loop_begin",
    "b2mpc_closed_loop",
]


def lib():
    """Load libb200mpc.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2mpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
    L.b2mpc_abi_version.restype = C.c_int
    L.b2mpc_last_error.restype = C.c_char_p
    L.b2mpc_default_config.argtypes = [C.POINTER(Config), i32, i32, C.c_double, i32, i32]
    L.b2mpc_default_config.restype = None
    L.b2mpc_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.b2mpc_destroy.argtypes = [vp]
    L.b2mpc_field_len.argtypes = [vp, C.c_int]; L.b2mpc_field_len.restype = i64
    L.b2mpc_workspace_bytes.argtypes = [vp]; L.b2mpc_workspace_bytes.restype = i64
    L.b2mpc_launch_count.argtypes = [vp]; L.b2mpc_launch_count.restype = i64
    L.b2mpc_kernel.argtypes = [vp]
    L.b2mpc_field_ptr.argtypes = [vp, C.c_int]; L.b2mpc_field_ptr.restype = vp
    for name in ("b2mpc_set", "b2mpc_get"):
        getattr(L, name).argtypes = [vp, C.c_int, vp, i64, C.c_int, vp]
    L.b2mpc_get_i32.argtypes = [vp, C.c_int, vp, i64, vp]
    for name in ("b2mpc_set_host", "b2mpc_get_host", "b2mpc_get_i32_host"):
        getattr(L, name).argtypes = [vp, C.c_int, vp, i64, vp]
    L.b2mpc_reset.argtypes = [vp, vp, i64, vp]
    for name in ("b2mpc_prepare_slack", "b2mpc_solve", "b2mpc_sim", "b2mpc_shift"):
        getattr(L, name).argtypes = [vp, i64, vp]
    L.b2mpc_solve_sqp.argtypes = [vp, i64, i32, dp, vp]
    L.b2mpc_solve_host.argtypes = [vp, i64] + [vp] * 11 + [vp]
    L.b2mpc_dfma_peak.argtypes = [C.c_int, dp]
    L.b2mpc_obst_predict.argtypes = [vp, i64, i32, vp]
    L.b2mpc_loop_reset.argtypes = [vp, i64, vp]
    L.b2mpc_loop_begin.argtypes = [vp, i64, vp]
    L.b2mpc_closed_loop.argtypes = [vp, i64, i32, i32, C.c_uint64, i64, vp, i32, i32, i32, C.c_double, vp]
    L.b2mpc_loop_post.argtypes = [vp, i64, vp, i32, i32, C.c_double, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("b2mpc_last_error", "b2mpc_default_config", "b2mpc_field_len", "b2mpc_workspace_bytes",
                        "b2mpc_launch_count", "b2mpc_field_ptr"):
            fn.restype = C.c_int
    if L.b2mpc_abi_version() != 3:
        raise B2mpcError("libb200mpc ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise B2mpcError(f"libb200mpc error {rc}: {lib().b2mpc_last_error().decode()}")


def default_config(N=20, n_obst=5, tf=2.0, max_batch=4096, device=0, **overrides) -> Config:
    c = Config()
    lib().b2mpc_default_config(C.byref(c), N, n_obst, tf, max_batch, device)
    for k, v in overrides.items():
        cur = getattr(c, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(c, k, v)
    return c
