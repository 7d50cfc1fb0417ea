#!/usr/bin/env bash
# Build libretisar_b200.so for sm_100a (B200) in-tree.  Usage: retisar_b200/csrc/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${RTB_OUT_DIR:-$here/../lib}"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC -shared \
    -Xptxas -v "$@" \
    -o "$out/libretisar_b200.so" \
    "$here/capi.cu" "$here/sh_plan.cu" "$here/ols_plan.cu" "$here/fd_plan.cu" "$here/output_stage.cu" \
    -cudart static
echo "built $out/libretisar_b200.so"
