"""Batched solver object: the B200 counterpart of the `AcadosOcpSolver` + `AcadosSimSolver` pair the
reference builds at src/simulation/robot_ocp_problem.py:135-136, for B independent problems.

PyTorch is plumbing only (device tensors, streams); all arithmetic runs inside libb200mpc.so.
Tensors at this interface are batch-first (`[B, N+1, 5]` ...), float64; CUDA tensors are converted to
the library's batch-last layout on device, CPU tensors / numpy arrays go through the `*_host` entry
points (H2D/D2H inside the call).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class BatchedMpcSolver:
    """set / cost_set / solve / get for a batch of OCPs (robot_ocp_problem.py:140-258)."""

    def __init__(self, N=20, n_obst=5, tf=2.0, max_batch=4096, device=0, **overrides):
        import torch
        if not torch.cuda.is_available():
            raise capi.B2mpcError("no CUDA device visible: the B200 solver has no CPU fallback")
        self.torch = torch
        self.cfg = capi.default_config(N, n_obst, tf, max_batch, device, **overrides)
        self.N, self.M, self.tf, self.max_batch, self.device = N, n_obst, tf, max_batch, device
        self._h = C.c_void_p()
        capi.check(capi.lib().b2mpc_create(C.byref(self.cfg), C.byref(self._h)))
        self.dt = tf / N
        self._shapes = {
            capi.X0: (5,), capi.P: (N + 1, 2 * n_obst), capi.YREF: (N + 1, 6), capi.ZL: (N + 1, n_obst),
            capi.ZZL: (N + 1, n_obst), capi.X: (N + 1, 5), capi.U: (N, 2), capi.SL: (N + 1, n_obst),
            capi.SU: (N + 1, n_obst), capi.PI: (N, 5), capi.LAM: (N + 1, 12 + 4 * n_obst),
            capi.T: (N + 1, 12 + 4 * n_obst), capi.STATS: (8,), capi.GOAL: (2,), capi.OBST: (n_obst, 4),
            capi.MARGIN: (2,), capi.U0: (2,), capi.NLP_RES: (4,),
        }

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if self._h:
            capi.lib().b2mpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _fid(self, field):
        return capi.FIELD_NAMES[field] if isinstance(field, str) else field

    # -- data movement ----------------------------------------------------------------------------
    def set(self, field, value):
        """value: [B, ...] float64; torch CUDA tensor (device path) or numpy / CPU tensor (host path)."""
        fid = self._fid(field)
        torch = self.torch
        if isinstance(value, torch.Tensor) and value.is_cuda:
            v = value.to(torch.float64).contiguous()
            B = v.shape[0]
            assert tuple(v.shape[1:]) == self._shapes[fid], (v.shape, self._shapes[fid])
            capi.check(capi.lib().b2mpc_set(self._h, fid, C.c_void_p(v.data_ptr()), B, capi.LAYOUT_AOS, self._stream()))
            self._keep = v      # keep alive until the stream has consumed it
        else:
            a = value.numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
            a = np.ascontiguousarray(a, np.float64)
            B = a.shape[0]
            assert tuple(a.shape[1:]) == self._shapes[fid], (a.shape, self._shapes[fid])
            capi.check(capi.lib().b2mpc_set_host(self._h, fid, a.ctypes.data_as(C.c_void_p), B, self._stream()))
            self.torch.cuda.current_stream(self.device).synchronize()   # pageable source must outlive the copy
        return B

    def get(self, field, batch, host=False):
        fid = self._fid(field)
        shape = (batch,) + self._shapes[fid]
        if host:
            out = np.empty(shape, np.float64)
            capi.check(capi.lib().b2mpc_get_host(self._h, fid, out.ctypes.data_as(C.c_void_p), batch, self._stream()))
            return out
        out = self.torch.empty(shape, dtype=self.torch.float64, device=f"cuda:{self.device}")
        capi.check(capi.lib().b2mpc_get(self._h, fid, C.c_void_p(out.data_ptr()), batch, capi.LAYOUT_AOS, self._stream()))
        return out

    def get_int(self, ifield, batch, host=False):
        ifield = capi.IFIELD_NAMES.get(ifield, ifield)
        if host:
            out = np.empty(batch, np.int32)
            capi.check(capi.lib().b2mpc_get_i32_host(self._h, ifield, out.ctypes.data_as(C.c_void_p), batch, self._stream()))
            return out
        out = self.torch.empty(batch, dtype=self.torch.int32, device=f"cuda:{self.device}")
        capi.check(capi.lib().b2mpc_get_i32(self._h, ifield, C.c_void_p(out.data_ptr()), batch, self._stream()))
        return out

    # -- the reference's per-step calls ---------------------------------------------------------
    def prepare_slack(self, batch):
        """parameterize_slack(), robot_ocp_problem.py:145-152 (needs 'x0' and 'goal' set)."""
        capi.check(capi.lib().b2mpc_prepare_slack(self._h, batch, self._stream()))

    def reset(self, batch, mask=None):
        """ocp_solver.reset() + set_initial_guess(), :286-306."""
        mp = None
        if mask is not None:
            mask = mask.to(self.torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        capi.check(capi.lib().b2mpc_reset(self._h, mp, batch, self._stream()))

    def solve(self, batch):
        """ocp_solver.solve(), :195 -- asynchronous on the current stream."""
        capi.check(capi.lib().b2mpc_solve(self._h, batch, self._stream()))

    def solve_sqp(self, batch, max_iter=0, tol=None):
        """acados' full `SQP` (nlp_solver_type 'SQP': residual test -> QP -> full step, repeated) for every problem in
        ONE launch; max_iter = nlp_solver_max_iter, tol = (tol_stat, tol_eq, tol_ineq, tol_comp); 0 / None take the
        handle's configuration.  Status 0 converged, 2 max_iter, 4 QP failure; 'sqp_iter', 'nlp_res' hold the
        statistics, 'qp_iter' the interior-point iterations in total.  The reference itself runs SQP_RTI (:130)."""
        tp = None
        if tol is not None:
            tp = (C.c_double * 4)(*[float(v) for v in tol])
        capi.check(capi.lib().b2mpc_solve_sqp(self._h, batch, int(max_iter), tp, self._stream()))

    def reset_multipliers(self, batch):
        """Forget pi / lam / t (what ocp_solver.reset() does to the multipliers) without touching the iterate: the
        full-SQP mode's first convergence test reads them."""
        z = self.torch.zeros
        dev = f"cuda:{self.device}"
        for f in (capi.PI, capi.LAM, capi.T):
            self.set(f, z((batch,) + self._shapes[f], dtype=self.torch.float64, device=dev))

    def sim(self, batch):
        """ocp_integrator.solve(), :207-212."""
        capi.check(capi.lib().b2mpc_sim(self._h, batch, self._stream()))

    def shift(self, batch):
        """warm-start shift, :253-258."""
        capi.check(capi.lib().b2mpc_shift(self._h, batch, self._stream()))

    # -- closed loop around the solve (robot_ocp_problem.py:184-260) ------------------------------
    def obst_predict(self, batch, quirk=True):
        """parameterize_model(), :154-166: 'p' <- horizon prediction of 'obst' (visualization.py:62-79)."""
        capi.check(capi.lib().b2mpc_obst_predict(self._h, batch, 1 if quirk else 0, self._stream()))

    def loop_reset(self, batch):
        capi.check(capi.lib().b2mpc_loop_reset(self._h, batch, self._stream()))

    def loop_begin(self, batch):
        """head of RobotOcpProblem.step() (:177-183): clears done / out-of-bounds / step counter, keeps the episode's min margin"""
        capi.check(capi.lib().b2mpc_loop_begin(self._h, batch, self._stream()))

    def loop_post(self, batch, noise=None, random_move=False, init_guess_when_error=True, tol=0.15):
        """everything of one control step after solve(), :198-260; noise: CUDA tensor [B,M,2] ~ N(0,1)."""
        npnt = None
        if noise is not None:
            noise = noise.to(self.torch.float64).contiguous()
            assert noise.is_cuda and tuple(noise.shape) == (batch, self.M, 2)
            npnt = C.c_void_p(noise.data_ptr())
            self._keep_noise = noise
        capi.check(capi.lib().b2mpc_loop_post(self._h, batch, npnt, 1 if random_move else 0,
                                              1 if init_guess_when_error else 0, float(tol), self._stream()))

    def closed_loop(self, batch, steps, noise=None, seed=0, env_offset=0, random_move=False, init_guess_when_error=True,
                    quirk=True, tol=0.15):
        """`steps` control steps of robot_ocp_problem.py:184-260 for every environment, enqueued by one C call
        (b2mpc_closed_loop).  noise: CUDA tensor [steps,B,M,2] ~ N(0,1) to replay; without it (and random_move) the
        on-device Philox stream keyed by (seed, env_offset + environment, control step, obstacle) is used."""
        mode, npnt = 0, None
        if random_move:
            if noise is not None:
                noise = noise.to(self.torch.float64).contiguous()
                assert noise.is_cuda and tuple(noise.shape) == (steps, batch, self.M, 2), tuple(noise.shape)
                mode, npnt = 2, C.c_void_p(noise.data_ptr())
                self._keep_noise = noise
            else:
                mode = 1
        capi.check(capi.lib().b2mpc_closed_loop(self._h, batch, steps, mode, int(seed) & (2 ** 64 - 1), int(env_offset), npnt,
                                                1 if random_move else 0, 1 if init_guess_when_error else 0,
                                                1 if quirk else 0, float(tol), self._stream()))

    def solve_host(self, x0, p, yref, zl, Zl, x, u, sl, su):
        """One control step on HOST numpy buffers (batch-first): copies in, solves, copies the iterate
        back IN PLACE; returns (status, qp_iter)."""
        B = x0.shape[0]
        arrs = [x0, p, yref, zl, Zl, x, u, sl, su]
        for a in arrs:
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape[0] == B
        status = np.empty(B, np.int32); qp_iter = np.empty(B, np.int32)
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in arrs]
        capi.check(capi.lib().b2mpc_solve_host(self._h, B, *ptrs, status.ctypes.data_as(C.c_void_p),
                                               qp_iter.ctypes.data_as(C.c_void_p), self._stream()))
        return status, qp_iter

    @property
    def kernel_id(self):
        """2 = warp-per-problem with the node-private rows in global memory (persistent grid), 3 = everything in shared memory"""
        return capi.lib().b2mpc_kernel(self._h)

    @property
    def kernel_name(self):
        return {2: "warp-per-problem", 3: "warp-per-problem (all shared)", 4: "cta-per-problem"}[self.kernel_id]

    @property
    def launch_count(self):
        return capi.lib().b2mpc_launch_count(self._h)

    @property
    def workspace_bytes(self):
        return capi.lib().b2mpc_workspace_bytes(self._h)


def latency_kernel(N, n_obst, nlp_solver=0):
    """`kernel` value for ONE problem at a time (the reference's own use: a single closed-loop episode): the CTA-per-problem
    kernel where it is instantiated (5 obstacles, N = 20 or 10, SQP_RTI), else 0 (auto).  Batched throughput runs keep the
    default: its results do not depend on the batch size or on the sharding, bit for bit."""
    return 4 if (n_obst == 5 and N in (20, 10) and nlp_solver == 0) else 0


def dfma_peak_tflops(device=0):
    out = C.c_double(0.0)
    capi.check(capi.lib().b2mpc_dfma_peak(device, C.byref(out)))
    return out.value
