#!/usr/bin/env bash
# One GPU visit: smoke, GPU parity tests, a short bench.  Logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
for v in auto shift; do
timeout 600 python bench.py --steps 10 --warmup 3 --variant $v ${BENCH_EXTRA} > gpurun_out/bench_$v.log 2>&1; echo "bench $v rc=$?" | tee -a gpurun_out/bench_$v.log
tail -2 gpurun_out/bench_$v.log
done
