#!/usr/bin/env bash
# Source-level ncu capture of ONE launch (warp-stall samples per SASS line): SKIP = launch index among conv_mma
# launches of `bench.py --steps 1 --warmup 1` (54 = layer 0 of the timed step), TAG
mkdir -p gpurun_out
T=${TAG:-s}
for sk in ${SKIPS:-58}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-conv_mma} -s $sk -c 1 -o gpurun_out/src_${T}_$sk -f python bench.py ${BENCH_EXTRA} --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_src_${T}_$sk.log 2>&1; echo "ncu rc=$?"
  python tools/ncu_sass_hist.py gpurun_out/src_${T}_$sk.ncu-rep ::regex:${KERNEL:-conv_mma}:1 60 > gpurun_out/sass_${T}_$sk.txt 2>&1
  python tools/ncu_stalls.py gpurun_out/src_${T}_$sk.ncu-rep > gpurun_out/stalls_${T}_$sk.txt 2>&1
  rm -f gpurun_out/src_${T}_$sk.ncu-rep
done
