"""In-tree build of libb200mpc.so with nvcc for sm_100a (B200).  No JIT cache, no pip install: the
.so lands in dynamic-obstacle-avoidance-mpc_b200/lib/ and travels with the repo snapshot."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "lib", "libb200mpc.so")
# -fmad=false: no implicit floating-point contraction.  Every fused multiply-add of the kernels is written as fma() in the
# source, so that the CPU mirror build (oracle/simt/, g++ -ffp-contract=off) executes the same IEEE-754 operations in the
# same order and the parity tests can demand bit-identical results.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return [os.path.join(SRC, f) for f in sorted(os.listdir(SRC)) if f.endswith((".cu", ".cuh", ".h")) and f != "model_host.cu"] + \
           [os.path.join(os.path.dirname(PKG), "include", "b200mpc.h")]


MODEL_HOST_OUT = os.path.join(PKG, "lib", "libb200model_host.so")


def build_model_host(force=False):
    """host build of the generated model functions (CPU tests of the generated code)"""
    src = os.path.join(SRC, "model_host.cu")
    deps = [src, os.path.join(SRC, "model_generated.cuh"), os.path.join(SRC, "b2_math.cuh")]
    os.makedirs(os.path.dirname(MODEL_HOST_OUT), exist_ok=True)
    stale = (not os.path.exists(MODEL_HOST_OUT)) or any(os.path.getmtime(s) > os.path.getmtime(MODEL_HOST_OUT) for s in deps)
    if force or stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", MODEL_HOST_OUT, src])
    return MODEL_HOST_OUT


def build(force=False, verbose=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stale = (not os.path.exists(OUT)) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in sources())
    if not (force or stale):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, os.path.join(SRC, "b200mpc.cu")]
    subprocess.check_call(cmd)
    return OUT
