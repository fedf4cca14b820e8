#!/usr/bin/env bash
# Builds the compiled-reference pieces of the oracle into oracle/_ref/ (git-ignored).
# Reference sources are compiled where they lie under $TF2_REFERENCE (default /root/reference);
# nothing is copied into the repo.  Safe to run when the reference is absent (does nothing).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${TF2_REFERENCE:-/root/reference}"
out="$here/_ref"
[ -d "$ref/Runtime_Engine/cnn/host/src" ] || { echo "reference not present at $ref; skipping"; exit 0; }
mkdir -p "$out"
host="$ref/Runtime_Engine/cnn/host"
common="$ref/Runtime_Engine/common/inc"
for net in RESNET50 GOOGLENET RESNET50_PRUNED; do
  lower=$(echo "$net" | tr 'A-Z' 'a-z')
  g++ -std=c++11 -O1 -fPIC -shared -w -fopenmp -D"$net" -DPRINT_LEVEL_QUIET \
      -I"$here/stub" -I"$common" -I"$host/inc" \
      "$host/src/model_loader.cpp" "$host/src/quantization.cpp" "$host/src/input_loader.cpp" \
      "$host/src/debug.cpp" "$host/src/network_helper.cpp" "$here/ref_host_shim.cpp" \
      -o "$out/libtf2ref_host_${lower}.so"
done
echo "built: $(ls "$out")"
# the reference PE kernel (device/src/pe.cl) compiled as C behind the FIFO shim of ref_device/pe_harness.c
dev="$ref/Runtime_Engine/cnn/device/src"
/usr/bin/gcc -x c -std=gnu11 -O1 -fPIC -shared -w -DRESNET50 -I"$dev" -I"$host/inc" \
    "$here/ref_device/pe_harness.c" -o "$out/libtf2ref_pe.so"
echo "built: libtf2ref_pe.so"
# the reference's post-PE kernels (relu / pool / pool_tail / feature_writer / full_size_pool) compiled as
# C with each network's own tables (ref_device/post_harness.c)
for net in RESNET50 GOOGLENET RESNET50_PRUNED; do
  lower=$(echo "$net" | tr 'A-Z' 'a-z')
  /usr/bin/gcc -x c -std=gnu11 -O1 -fPIC -shared -w -D"$net" -I"$dev" -I"$host/inc" -I"$here/ref_device" \
      "$here/ref_device/post_harness.c" -o "$out/libtf2ref_post_${lower}.so"
done
echo "built: libtf2ref_post_{resnet50,googlenet,resnet50_pruned}.so"
# the reference's whole device pipeline (cnn.cl) for layer 0 of each network (ref_device/full_harness.c)
for net in RESNET50 GOOGLENET RESNET50_PRUNED; do
  lower=$(echo "$net" | tr 'A-Z' 'a-z')
  /usr/bin/gcc -x c -std=gnu11 -O1 -fPIC -shared -w -D"$net" -I"$dev" -I"$host/inc" -I"$here/ref_device" \
      "$here/ref_device/full_harness.c" -o "$out/libtf2ref_full_${lower}.so"
done
echo "built: libtf2ref_full_{resnet50,googlenet,resnet50_pruned}.so"
# the reference's whole device program over ALL layers, kernels as coroutines (ref_device/net_harness.c)
for net in RESNET50 GOOGLENET RESNET50_PRUNED; do
  lower=$(echo "$net" | tr 'A-Z' 'a-z')
  /usr/bin/gcc -x c -std=gnu11 -O1 -fPIC -shared -w -D"$net" -I"$dev" -I"$host/inc" -I"$here/ref_device" \
      "$here/ref_device/net_harness.c" -o "$out/libtf2ref_net_${lower}.so"
done
echo "built: libtf2ref_net_{resnet50,googlenet,resnet50_pruned}.so"
