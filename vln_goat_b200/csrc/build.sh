#!/bin/bash
# Builds vln_goat_b200/libgoat_sm100.so in-tree (sm_100a only).  nvcc cross-compiles without a GPU.
# Each .cu is compiled to build/*.o in parallel (only when it or a header changed), then linked.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libgoat_sm100.so"
OBJ="$HERE/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p "$OBJ"
pids=()
for src in "$HERE"/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/../../include/goat_sm100.h" -nt "$o" ] \
     || [ "$HERE/build.sh" -nt "$o" ]; then
    "$NVCC" $FLAGS "$@" -c "$src" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
