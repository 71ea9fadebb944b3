"""`acados_template`-compatible classes backed by libb200mpc (batch = 1 view of the batched solver).

The reference drives its solver exclusively through this surface:
  AcadosModel / AcadosOcp attribute bags                       robot_model.py:47-65, robot_ocp_problem.py:55-132
  AcadosOcpSolver(ocp, json_file=).set / cost_set / set_params_sparse / solve / get / reset / get_stats
                                                               robot_ocp_problem.py:135,140,151-152,165-166,191-198,232,254-258,291
  AcadosSimSolver(ocp, json_file=).set / solve / get           robot_ocp_problem.py:136,207-212

`install()` registers these classes (plus the SymPy `casadi` shim and an `imp` stub) in `sys.modules`,
after which `src/simulation/robot_ocp_problem.py` runs UNMODIFIED on the B200 path.  Values are
mirrored on the host like acados does (callers keep ownership of what they pass); the CUDA handle is
created on the first `solve()`, so constructing and parameterising the problem needs no GPU -- solving
does, and there is no CPU fallback.
"""
from __future__ import annotations

import os
import sys
import time
import types

import numpy as np

_IDXBX = [0, 1, 3, 4]


class _Bag:
    """attribute bag (acados_template uses plain classes with defaults)"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class AcadosModel:
    def __init__(self):
        self.name = None
        self.x = self.xdot = self.u = self.z = self.p = None
        self.f_impl_expr = self.f_expl_expr = None
        self.con_h_expr = self.con_h_expr_e = None


class AcadosOcp:
    def __init__(self):
        self.model = AcadosModel()
        self.dims = _Bag(N=None)
        self.cost = _Bag(cost_type="LINEAR_LS", cost_type_e="LINEAR_LS", Vx=None, Vu=None, yref=None, Vx_0=None, Vu_0=None,
                         yref_0=None, Vx_e=None, yref_e=None, W=None, W_0=None, W_e=None, zl=None, zl_e=None, Zl=None,
                         Zl_e=None, Zu=None, Zu_e=None, zu=None, zu_e=None)
        self.constraints = _Bag(x0=None, lbx=None, ubx=None, idxbx=None, lbu=None, ubu=None, idxbu=None, lh=None, uh=None,
                                lh_e=None, uh_e=None, Jsh=None, Jsh_e=None)
        self.solver_options = _Bag(qp_solver="PARTIAL_CONDENSING_HPIPM", hessian_approx="GAUSS_NEWTON",
                                   levenberg_marquardt=0.0, integrator_type="ERK", nlp_solver_type="SQP",
                                   qp_solver_iter_max=50, tf=None, nlp_solver_max_iter=100, nlp_solver_tol_stat=1e-6,
                                   nlp_solver_tol_eq=1e-6, nlp_solver_tol_ineq=1e-6, nlp_solver_tol_comp=1e-6,
                                   globalization="FIXED_STEP", alpha_min=0.05, alpha_reduction=0.7)
        self.parameter_values = None


class AcadosSim:
    def __init__(self):
        self.model = AcadosModel()
        self.solver_options = _Bag(T=None, integrator_type="ERK", num_stages=4, num_steps=1, newton_iter=3, collocation_type="GAUSS_LEGENDRE")


def _n(expr):
    """length of a model expression vector (casadi shim SX or list)"""
    if hasattr(expr, "size"):
        s = expr.size()
        return s[0] if isinstance(s, tuple) else int(s)
    return len(expr)


def ocp_to_config(ocp: AcadosOcp) -> dict:
    """Validate an AcadosOcp against what the accelerated path implements and translate it into
    b2mpc_config overrides.  Anything outside the reference's problem class raises, like acados would
    at code-generation time."""
    N = int(ocp.dims.N)
    o, c, k = ocp.solver_options, ocp.cost, ocp.constraints
    if o.nlp_solver_type not in ("SQP_RTI", "SQP"):
        raise NotImplementedError(f"nlp_solver_type {o.nlp_solver_type!r}: 'SQP_RTI' (robot_ocp_problem.py:130) or 'SQP'")
    glob = {"FIXED_STEP": 0, "MERIT_BACKTRACKING": 1}.get(getattr(o, "globalization", "FIXED_STEP"))
    if glob is None:
        raise NotImplementedError(f"globalization {o.globalization!r}")
    if o.hessian_approx != "GAUSS_NEWTON" or o.qp_solver != "PARTIAL_CONDENSING_HPIPM":
        raise NotImplementedError("only GAUSS_NEWTON + PARTIAL_CONDENSING_HPIPM (robot_ocp_problem.py:126-127)")
    if c.cost_type != "LINEAR_LS" or c.cost_type_e != "LINEAR_LS":
        raise NotImplementedError("only LINEAR_LS costs (robot_ocp_problem.py:59-60)")
    integ = {"IRK": 0, "ERK": 1}.get(o.integrator_type)
    if integ is None:
        raise NotImplementedError(f"integrator_type {o.integrator_type!r}")
    nx, nu = _n(ocp.model.x), _n(ocp.model.u)
    if (nx, nu) != (5, 2):
        raise NotImplementedError("the generated device model has nx=5, nu=2 (src/models/robot_model.py)")
    M = _n(ocp.model.con_h_expr)
    if _n(ocp.model.p) != 2 * M:
        raise ValueError("parameter vector must hold one (x,y) pair per obstacle constraint")
    sel = np.zeros((6, 5)); sel[0, 0] = sel[1, 1] = sel[2, 3] = sel[3, 4] = 1
    selu = np.zeros((6, 2)); selu[4, 0] = selu[5, 1] = 1
    if not (np.array_equal(np.asarray(c.Vx), sel) and np.array_equal(np.asarray(c.Vu), selu) and np.array_equal(np.asarray(c.Vx_e), sel[:4])):
        raise NotImplementedError("Vx/Vu/Vx_e must select y=[x,y,v,omega,u_a,u_alpha] (robot_ocp_problem.py:63-76)")
    W, We = np.asarray(c.W, float), np.asarray(c.W_e, float)
    if np.count_nonzero(W - np.diag(np.diag(W))) or np.count_nonzero(We - np.diag(np.diag(We))):
        raise NotImplementedError("W, W_e must be diagonal")
    if list(np.asarray(k.idxbx).astype(int)) != _IDXBX or list(np.asarray(k.idxbu).astype(int)) != [0, 1]:
        raise NotImplementedError("idxbx must be [0,1,3,4] and idxbu [0,1] (robot_ocp_problem.py:93,97)")
    if k.Jsh is None or not np.array_equal(np.asarray(k.Jsh), np.eye(M)):
        raise NotImplementedError("all obstacle constraints must be softened (Jsh = I, robot_ocp_problem.py:109)")
    lh, uh = np.asarray(k.lh, float), np.asarray(k.uh, float)
    if np.ptp(lh) or np.ptp(uh) or not np.array_equal(lh, np.asarray(k.lh_e, float)) or not np.array_equal(uh, np.asarray(k.uh_e, float)):
        raise NotImplementedError("lh/uh must be uniform and equal to lh_e/uh_e")
    zu, Zu = np.asarray(c.zu, float), np.asarray(c.Zu, float)
    if np.ptp(zu) or np.ptp(Zu):
        raise NotImplementedError("zu/Zu must be uniform")
    return dict(N=N, n_obst=M, tf=float(o.tf), integrator=integ, qp_iter_max=int(o.qp_solver_iter_max),
                lm=float(o.levenberg_marquardt), W=list(np.diag(W)), W_e=list(np.diag(We)),
                lbx=list(np.asarray(k.lbx, float)), ubx=list(np.asarray(k.ubx, float)),
                lbu=list(np.asarray(k.lbu, float)), ubu=list(np.asarray(k.ubu, float)),
                lh=float(lh[0]), uh=float(uh[0]), zu=float(zu[0]), Zu=float(Zu[0]),
                nlp_solver=1 if o.nlp_solver_type == "SQP" else 0, nlp_iter_max=int(getattr(o, "nlp_solver_max_iter", 100)),
                nlp_tol=[float(getattr(o, "nlp_solver_tol_" + n, 1e-6)) for n in ("stat", "eq", "ineq", "comp")],
                nlp_globalization=glob, ls_alpha_min=float(getattr(o, "alpha_min", 0.05)),
                ls_alpha_reduction=float(getattr(o, "alpha_reduction", 0.7)))


class AcadosOcpSolver:
    """batch-1 facade over BatchedMpcSolver with acados' copy-in/copy-out semantics"""

    def __init__(self, acados_ocp: AcadosOcp, json_file="acados_ocp.json", device=0, **_ignored):
        self.acados_ocp = acados_ocp
        self._cfg = ocp_to_config(acados_ocp)
        self.N, self.M = self._cfg["N"], self._cfg["n_obst"]
        N, M = self.N, self.M
        self._device = device
        self._solver = None
        c, k = acados_ocp.cost, acados_ocp.constraints
        self._f = dict(
            x0=np.asarray(k.x0, float).reshape(1, 5).copy(),
            p=np.tile(np.asarray(acados_ocp.parameter_values, float).reshape(1, 1, 2 * M), (1, N + 1, 1)),
            yref=np.zeros((1, N + 1, 6)), zl=np.zeros((1, N + 1, M)), Zl=np.zeros((1, N + 1, M)),
            x=np.zeros((1, N + 1, 5)), u=np.zeros((1, N, 2)), sl=np.zeros((1, N + 1, M)), su=np.zeros((1, N + 1, M)))
        self._f["yref"][0, :N] = np.asarray(c.yref, float)
        self._f["yref"][0, N, :4] = np.asarray(c.yref_e, float)
        self._f["zl"][0, :N] = np.asarray(c.zl, float); self._f["zl"][0, N] = np.asarray(c.zl_e, float)
        self._f["Zl"][0, :N] = np.asarray(c.Zl, float); self._f["Zl"][0, N] = np.asarray(c.Zl_e, float)
        self._stats = dict(time_tot=0.0, qp_iter=0, sqp_iter=1, status=0)

    # -- acados API ------------------------------------------------------------------------------------
    def set(self, stage, field, value):
        v = np.asarray(value, float).ravel()
        N = self.N
        if not 0 <= stage <= N:
            raise Exception(f"AcadosOcpSolver.set(): stage {stage} out of range")
        if field == "x":
            self._chk(v, 5, field); self._f["x"][0, stage] = v
        elif field == "u":
            if stage >= N:
                raise Exception("AcadosOcpSolver.set(): no control at the terminal node")
            self._chk(v, 2, field); self._f["u"][0, stage] = v
        elif field == "p":
            self._chk(v, 2 * self.M, field); self._f["p"][0, stage] = v
        elif field in ("lbx", "ubx"):
            if stage != 0:
                raise NotImplementedError("state bounds are fixed at construction except the x0 embedding at stage 0")
            self._chk(v, 5, field)
            # robot_ocp_problem.py:191-192 sets ubx then lbx to the same x0
            self._f["x0"][0] = v
        elif field == "sl":
            self._chk(v, self.M, field); self._f["sl"][0, stage] = v
        elif field == "su":
            self._chk(v, self.M, field); self._f["su"][0, stage] = v
        else:
            raise Exception(f"AcadosOcpSolver.set(): field {field!r} not supported by the B200 path")

    def cost_set(self, stage, field, value):
        v = np.asarray(value, float).ravel()
        if field == "zl":
            self._chk(v, self.M, field); self._f["zl"][0, stage] = v
        elif field == "Zl":
            self._chk(v, self.M, field); self._f["Zl"][0, stage] = v
        elif field == "yref":
            n = 6 if stage < self.N else 4
            self._chk(v, n, field)          # the reference's set_subgoal passes 5 values here (:284) -> raises, as in acados
            self._f["yref"][0, stage, :n] = v
        else:
            raise Exception(f"AcadosOcpSolver.cost_set(): field {field!r} not supported by the B200 path")

    def set_params_sparse(self, stage, idx, value):
        idx = np.asarray(idx, int).ravel(); v = np.asarray(value, float).ravel()
        if idx.size != v.size or (idx.size and (idx.min() < 0 or idx.max() >= 2 * self.M)):
            raise Exception("AcadosOcpSolver.set_params_sparse(): bad index / value")
        self._f["p"][0, stage, idx] = v

    def solve(self):
        if self._solver is None:
            from ..solver import BatchedMpcSolver, latency_kernel
            cfg = dict(self._cfg); N = cfg.pop("N"); M = cfg.pop("n_obst"); tf = cfg.pop("tf")
            cfg.setdefault("kernel", latency_kernel(N, M, cfg.get("nlp_solver", 0)))      # one problem at a time: the latency kernel
            self._solver = BatchedMpcSolver(N, M, tf, max_batch=1, device=self._device, **cfg)
        f = self._f
        t0 = time.perf_counter()
        status, qp_iter = self._solver.solve_host(f["x0"], f["p"], f["yref"], f["zl"], f["Zl"], f["x"], f["u"], f["sl"], f["su"])
        self._stats.update(time_tot=time.perf_counter() - t0, qp_iter=int(qp_iter[0]), status=int(status[0]))
        if self._cfg["nlp_solver"] == 1:
            self._stats["sqp_iter"] = int(self._solver.get_int("sqp_iter", 1, host=True)[0])
        return int(status[0])

    def get(self, stage, field):
        if field in ("x", "u", "sl", "su", "p"):
            return self._f[field][0, stage].copy()
        if field in ("pi", "lam", "t"):
            if self._solver is None:
                raise Exception("AcadosOcpSolver.get(): solve() first")
            return self._solver.get(field, 1, host=True)[0, stage].copy()
        raise Exception(f"AcadosOcpSolver.get(): field {field!r} not supported by the B200 path")

    def reset(self):
        for k in ("x", "u", "sl", "su"):
            self._f[k][:] = 0.0
        if self._solver is not None:
            self._solver.reset_multipliers(1)          # only the 'SQP' convergence test ever reads them

    def get_residuals(self):
        """[res_stat, res_eq, res_ineq, res_comp] of the last 'SQP' solve (acados: AcadosOcpSolver.get_residuals)"""
        if self._solver is None or self._cfg["nlp_solver"] != 1:
            raise Exception("AcadosOcpSolver.get_residuals(): needs nlp_solver_type 'SQP' and a solve()")
        return self._solver.get("nlp_res", 1, host=True)[0].copy()

    def get_stats(self, field):
        if field in ("time_tot", "qp_iter", "sqp_iter", "status"):
            return self._stats[field]
        raise Exception(f"AcadosOcpSolver.get_stats(): field {field!r} not supported by the B200 path")

    def get_status(self):
        return self._stats["status"]

    @staticmethod
    def _chk(v, n, field):
        if v.size != n:
            raise Exception(f"AcadosOcpSolver: dimension mismatch for {field!r}: got {v.size}, expected {n}")


class AcadosSimSolver:
    """ocp_integrator of robot_ocp_problem.py:136,207-212: one step of the OCP's integrator (dt = tf/N)"""

    def __init__(self, acados_sim_or_ocp, json_file="acados_sim.json", device=0, **_ignored):
        o = acados_sim_or_ocp.solver_options
        if isinstance(acados_sim_or_ocp, AcadosSim):
            self._dt = float(o.T); self._N = 1
        else:
            self._dt = float(o.tf) / int(acados_sim_or_ocp.dims.N); self._N = int(acados_sim_or_ocp.dims.N)
        self._integ = {"IRK": 0, "ERK": 1}[o.integrator_type]
        self._M = max(1, _n(acados_sim_or_ocp.model.con_h_expr)) if acados_sim_or_ocp.model.con_h_expr is not None else 1
        self._x = np.zeros((1, 5)); self._u = np.zeros((1, 1, 2)); self._device = device; self._solver = None

    def set(self, field, value):
        v = np.asarray(value, float).ravel()
        if field == "x":
            self._x[0] = v
        elif field == "u":
            self._u[0, 0] = v
        elif field in ("xdot", "T"):
            pass
        else:
            raise Exception(f"AcadosSimSolver.set(): field {field!r} not supported")

    def solve(self):
        if self._solver is None:
            from ..solver import BatchedMpcSolver
            self._solver = BatchedMpcSolver(1, self._M, self._dt, max_batch=1, device=self._device, integrator=self._integ)
        self._solver.set("x0", self._x); self._solver.set("u", self._u)
        self._solver.sim(1)
        self._x = self._solver.get("x0", 1, host=True)
        return 0

    def get(self, field):
        if field == "x":
            return self._x[0].copy()
        raise Exception(f"AcadosSimSolver.get(): field {field!r} not supported")


def install():
    """Make `import acados_template`, `import casadi`, `import imp` resolve to the B200-backed stand-ins."""
    here = os.path.dirname(os.path.abspath(__file__))
    shims = os.path.join(os.path.dirname(os.path.dirname(here)), "codegen", "shims")
    me = sys.modules[__name__]
    pkg = types.ModuleType("acados_template")
    for name in ("AcadosModel", "AcadosOcp", "AcadosOcpSolver", "AcadosSimSolver", "AcadosSim"):
        setattr(pkg, name, getattr(me, name))
    sys.modules["acados_template"] = pkg
    for sub, names in (("acados_model", ["AcadosModel"]), ("acados_ocp", ["AcadosOcp"]), ("acados_ocp_solver", ["AcadosOcpSolver"]),
                       ("acados_sim_solver", ["AcadosSimSolver"]), ("acados_sim", ["AcadosSim"])):
        m = types.ModuleType("acados_template." + sub)
        for n in names:
            setattr(m, n, getattr(me, n))
        sys.modules["acados_template." + sub] = m
        setattr(pkg, sub, m)
    if "casadi" not in sys.modules:
        import importlib.util
        for modname in ("casadi", "imp"):
            if modname == "imp" and "imp" in sys.modules:
                continue
            spec = importlib.util.spec_from_file_location(modname, os.path.join(shims, modname + ".py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sys.modules[modname] = mod
    return pkg
