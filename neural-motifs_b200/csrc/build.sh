#!/usr/bin/env bash
# Builds libmotifs_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p build
objs=()
pids=()
for f in *.cu; do
  o="build/${f%.cu}.o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o" 2>/dev/null)" ]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "build/${f%.cu}.log" 2>&1 || { cat "build/${f%.cu}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o libmotifs_b200.so "${objs[@]}"
echo "built $(pwd)/libmotifs_b200.so"
