#!/usr/bin/env bash
# Round-2: step time against batch size (fixed cost per step = intercept of the line)
mkdir -p gpurun_out
for b in 64 128 256 512; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_batch_$b.log 2>&1
  tail -1 gpurun_out/bench_batch_$b.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('batch $b', round(l['value']), l['ms_per_step'], round(l['e2e']['value']))"
done
