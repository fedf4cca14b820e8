#!/usr/bin/env bash
# A/B visit: GPU parity tests + bench with the new kernel modes; on failure isolate by switch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -15 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then
  TF2B_MMA_HALO=0 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_nohalo.log 2>&1; echo "pytest (HALO=0) rc=$?"; tail -3 gpurun_out/pytest_nohalo.log
  TF2B_MMA_FOLD=0 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_nofold.log 2>&1; echo "pytest (FOLD=0) rc=$?"; tail -3 gpurun_out/pytest_nofold.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_auto.json > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_auto.log | cut -c1-200
TF2B_MMA_HALO=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_nohalo.json > gpurun_out/bench_nohalo.log 2>&1; echo "bench HALO=0 rc=$?"; tail -1 gpurun_out/bench_nohalo.log | cut -c1-200
TF2B_MMA_FOLD=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_nofold.json > gpurun_out/bench_nofold.log 2>&1; echo "bench FOLD=0 rc=$?"; tail -1 gpurun_out/bench_nofold.log | cut -c1-200
