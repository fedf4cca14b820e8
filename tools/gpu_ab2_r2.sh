#!/usr/bin/env bash
# A/B of run-time options on one box: bench lines with different executor / stem / weight-staging settings
mkdir -p gpurun_out
i=0
while read -r name flags; do
  [ -z "$name" ] && continue
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $flags > gpurun_out/bench_opt_$name.log 2>&1; echo "bench $name rc=$?"
  tail -1 gpurun_out/bench_opt_$name.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$name', round(l['value']), round(l['ms_per_step'],4), round(l['e2e']['value']), l['gpu_launches']//20)"
done <<< "${CASES}"
