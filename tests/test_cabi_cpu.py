"""CPU-side checks of the C-ABI (no GPU in this container): the shared library loads, exports every
function declared in include/b200mpc.h, fills the reference defaults, validates arguments, and REFUSES to
run without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from b200mpc import capi, build_ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_ext.build()
    return capi.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(b2mpc_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/b200mpc.h but not exported"
    assert set(capi.EXPORTS) <= names
    assert lib.b2mpc_abi_version() == 3


def test_default_config_matches_reference_constants(lib):
    c = capi.default_config(20, 5, 2.0, 4096, 0)
    assert (c.N, c.n_obst, c.max_batch, c.integrator, c.qp_iter_max) == (20, 5, 4096, 0, 50)
    assert list(c.W) == [2, 2, 2, 2, 0.15, 0.15] and list(c.W_e) == [5, 5, 5, 5]          # robot_ocp_problem.py:24-27
    assert list(c.lbx) == [-7, -7, -10, -10] and list(c.ubx) == [7, 7, 10, 10]            # :91-92
    assert list(c.lbu) == [-8, -8] and list(c.ubu) == [8, 8] and c.lm == 2.0              # :95-96,128
    assert abs(c.r2 - 2.4 ** 2) < 1e-15 and c.lh == 0.0 and c.uh == 1e5                   # robot_model.py:62, :99-103
    assert (c.mu0, c.alpha_min, c.res_g_max, c.res_b_max) == (1.0, 1e-8, 1e-6, 1e-8)      # acados' HPIPM block
    assert c.lm_scaled_by_dt == 1 and c.pred_corr == 1 and c.cond_pred_corr == 1 and c.itref_corr_max == 2
    assert c.nlp_solver == 0                                                              # 'SQP_RTI', :130
    assert c.nlp_iter_max == 100 and list(c.nlp_tol) == [1e-6] * 4                        # acados defaults (SQP mode)


def test_create_validates_arguments_and_has_no_cpu_fallback(lib):
    import torch
    h = C.c_void_p()
    for bad in (dict(N=0), dict(N=129), dict(n_obst=0), dict(n_obst=33), dict(max_batch=0), dict(tf=0.0), dict(integrator=7),
                dict(nlp_solver=2), dict(nlp_solver=1, nlp_iter_max=0), dict(kernel=1)):
        c = capi.default_config(20, 5, 2.0, 8, 0)
        for k, v in bad.items():
            setattr(c, k, v)
        assert lib.b2mpc_create(C.byref(c), C.byref(h)) == -1          # B2MPC_EINVAL, before touching CUDA
        assert lib.b2mpc_last_error()
    if not torch.cuda.is_available():
        c = capi.default_config(20, 5, 2.0, 8, 0)
        rc = lib.b2mpc_create(C.byref(c), C.byref(h))
        assert rc == -2 and b"no CUDA device" in lib.b2mpc_last_error()   # B2MPC_ENODEV: fails loudly
        from b200mpc import BatchedMpcSolver
        with pytest.raises(capi.B2mpcError):
            BatchedMpcSolver(20, 5, 2.0, max_batch=8)
        tf = C.c_double(0)
        assert lib.b2mpc_dfma_peak(0, C.byref(tf)) == -2
    assert lib.b2mpc_destroy(None) == 0
    assert lib.b2mpc_field_len(None, 0) == -1


def test_missing_library_is_a_loud_error(tmp_path, monkeypatch):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(capi.B2mpcError):
        capi.lib()


def test_ctypes_config_layout_matches_the_header(tmp_path):
    """the ctypes mirror of b2mpc_config (b200mpc/capi.py) against what a C compiler makes of include/b200mpc.h"""
    import subprocess
    names = [n for n, _ in capi.Config._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200mpc.h"\nint main(void){printf("%zu\\n", sizeof(b2mpc_config));\n' +
                   "".join(f'printf("%zu\\n", offsetof(b2mpc_config, {n}));\n' for n in names) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(capi.Config)
    for n, off in zip(names, out[1:]):
        assert getattr(capi.Config, n).offset == off, n
    # and the field / status enumerations
    hdr = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    for name, val in (("B2MPC_NLP_RES", capi.NLP_RES), ("B2MPC_N_FIELDS", capi.NLP_RES + 1), ("B2MPC_SQP_ITER", capi.SQP_ITER),
                      ("B2MPC_N_IFIELDS", capi.SQP_ITER + 1), ("B2MPC_U0", capi.U0), ("B2MPC_STATS", capi.STATS)):
        assert re.search(rf"\b{name}\s*=\s*{val}\b", hdr), name
