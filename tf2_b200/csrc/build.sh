#!/usr/bin/env bash
# Builds libtf2b200.so (sm_100a only) in-tree: tf2_b200/lib/libtf2b200.so
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="${TF2B_OUT:-$here/../lib}"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -cudart static)
objs=()
for f in conv_sa aux_kernels conv_mma api; do
  o="$out/$f.o"
  if [ ! -f "$o" ] || [ "$here/$f.cu" -nt "$o" ] || [ "$here/common.cuh" -nt "$o" ] || [ "$here/../../include/tf2b200.h" -nt "$o" ]; then
    "$NVCC" "${FLAGS[@]}" ${PTXAS_V:+-Xptxas -v} ${TF2B_EXPERIMENTS:+-DTF2B_EXPERIMENTS=1} ${TF2B_DEFS:-} -c "$here/$f.cu" -o "$o"
  fi
  objs+=("$o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$out/libtf2b200.so" "${objs[@]}"
echo "built $out/libtf2b200.so"
