#!/bin/sh
# Builds the C-ABI shared library for sm_100a (cross-compiles without a GPU).
# -fmad=false: the reference is built without FMA contraction; index/odds arithmetic must match bit for bit.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-O3,-Wall,-fvisibility=hidden -Xptxas -v -shared -cudart static \
  -o "${MLM_OUT:-$OUT/libmlmap_b200.so}" "$HERE/mlmap_capi.cu" "$@"
