#!/usr/bin/env bash
# Round-2: role counters + phase timeline of the (lean) producer / MMA loops, experiment library, one entry per distinct layer shape
mkdir -p gpurun_out
T=${TAG:-dbg}
export TF2B_LIB=$PWD/tools/micro/bin/exp_hs/libtf2b200.so
for n in ${NETS:-resnet50}; do
  TF2B_MMA_DEBUG=1 timeout 600 python bench.py --net $n --steps 1 --warmup 1 --executor 0 --no-cpu-baseline 2>&1 | grep -A1 "mma dbg" | grep -v "^--" | awk '/mma dbg/{k=$3 $4 $5 $6 $7; p=!seen[k]++} p' > gpurun_out/dbg_${T}_$n.txt
done
