#!/bin/bash
# experiment builds of libb200mpc.so with extra -D flags: scripts/build_variant.sh NAME [-DFLAG...]  -> lib/libb200mpc_NAME.so
# (used with B2MPC_LIB=dynamic-obstacle-avoidance-mpc_b200/lib/libb200mpc_NAME.so scripts/time_solve.py ...)
set -e
R=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
exec ${NVCC:-/usr/local/cuda/bin/nvcc} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" \
  -o "$R/dynamic-obstacle-avoidance-mpc_b200/lib/libb200mpc_$name.so" "$R/dynamic-obstacle-avoidance-mpc_b200/csrc/b200mpc.cu"
