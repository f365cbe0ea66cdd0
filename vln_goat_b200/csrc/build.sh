#!/bin/bash
# Builds vln_goat_b200/libgoat_sm100.so in-tree (sm_100a only).  nvcc cross-compiles without a GPU.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libgoat_sm100.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -o "$OUT" "$HERE"/*.cu "$@"
echo "built $OUT"
