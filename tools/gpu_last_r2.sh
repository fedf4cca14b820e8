#!/usr/bin/env bash
# Round-2 closing check: smoke, whole GPU suite, headline bench line
mkdir -p gpurun_out
T=${TAG:-last}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_resnet50_$T.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_resnet50_$T.log | cut -c1-200
