#!/usr/bin/env bash
# Round-2: halo tiles with streamed weights ("halo wstream") — parity subset, bench + per-layer table with the mode
# on (product library) and off / role counters (experiment library tools/micro/bin/exp_hs, TF2B_LIB override)
mkdir -p gpurun_out
T=${TAG:-hs}
timeout 1200 python -m pytest tests/test_gpu_mma.py tests/test_gpu_resnet50.py tests/test_vgg16.py tests/test_gpu_nets.py -m gpu -q ${PYTEST_EXTRA:--x} > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
for n in ${NETS:-resnet50 vgg16}; do
  timeout 600 python bench.py --net $n --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_${n}_$T.json > gpurun_out/bench_${n}_$T.log 2>&1; echo "bench $n rc=$?"
  tail -1 gpurun_out/bench_${n}_$T.log | cut -c1-200
done
export TF2B_LIB=$PWD/tools/micro/bin/exp_hs/libtf2b200.so
for hs in 1 0; do
  for n in ${NETS:-resnet50 vgg16}; do
    TF2B_MMA_HSTREAM=$hs timeout 600 python bench.py --net $n --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_${n}_${T}_exp$hs.json > gpurun_out/bench_${n}_${T}_exp$hs.log 2>&1; echo "bench exp hs=$hs $n rc=$?"
    tail -1 gpurun_out/bench_${n}_${T}_exp$hs.log | cut -c1-200
  done
  TF2B_MMA_HSTREAM=$hs TF2B_MMA_DEBUG=1 timeout 600 python bench.py --steps 1 --warmup 1 --executor 0 --no-cpu-baseline 2>&1 | grep "mma dbg" | grep " k3 " | awk '!seen[$3 $4 $5 $6 $7]++' > gpurun_out/dbg_${T}_hs$hs.txt
done
