#!/usr/bin/env bash
# Round-2: where do the K-heavy k x k layers lose time?  (1) parity subset on the product library, (2) experiment
# library: pipeline alone (TF2B_MMA_NOEPI=1) per-layer table, role counters + stage-position trace of the MMA warp
mkdir -p gpurun_out
T=${TAG:-hs2}
timeout 1200 python -m pytest tests/test_gpu_mma.py tests/test_gpu_resnet50.py tests/test_vgg16.py -m gpu -q ${PYTEST_EXTRA:--x} > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
export TF2B_LIB=$PWD/tools/micro/bin/exp_hs/libtf2b200.so
for hs in 1 0; do
  for ne in 0 1; do
    TF2B_MMA_HSTREAM=$hs TF2B_MMA_NOEPI=$ne timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_${T}_hs${hs}_ne$ne.json > gpurun_out/bench_${T}_hs${hs}_ne$ne.log 2>&1; echo "bench exp hs=$hs noepi=$ne rc=$?"
    tail -1 gpurun_out/bench_${T}_hs${hs}_ne$ne.log | cut -c1-160
  done
  TF2B_MMA_HSTREAM=$hs TF2B_MMA_DEBUG=1 TF2B_MMA_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --executor 0 --no-cpu-baseline 2>&1 | grep -A1 "mma dbg" | grep -A1 " k3 " | grep -v "^--" | awk '/mma dbg/{k=$3 $4 $5 $6 $7; p=!seen[k]++} p' > gpurun_out/dbg_${T}_hs$hs.txt
done
