#!/usr/bin/env bash
# Round-2 closing check: smoke, whole GPU suite, bench lines, executor modes on the branchy net
mkdir -p gpurun_out
T=${TAG:-last}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_resnet50_$T.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_resnet50_$T.log | cut -c1-200
for ex in 0 1 2 3; do
  timeout 600 python bench.py --net googlenet --executor $ex --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_googlenet_ex${ex}_$T.log 2>&1
  tail -1 gpurun_out/bench_googlenet_ex${ex}_$T.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('googlenet executor $ex', round(l['value']), l['ms_per_step'], round(l['e2e']['value']))"
done
