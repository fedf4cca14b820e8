#!/usr/bin/env bash
# Round-2 closing check: smoke, whole GPU suite, headline bench line (+ NETS)
mkdir -p gpurun_out
T=${TAG:-last}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
for n in ${NETS:-resnet50}; do
  timeout 600 python bench.py --net $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}_$T.log 2>&1; echo "bench $n rc=$?"; tail -1 gpurun_out/bench_${n}_$T.log | cut -c1-160
done
