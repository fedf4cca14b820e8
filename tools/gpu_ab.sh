#!/usr/bin/env bash
# A/B visit: GPU parity tests + bench under several experiment switches (VARIANTS="name:ENV=val,ENV=val ...")
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS}" ]; then
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_gpu.log
fi
for v in base ${VARIANTS}; do
  name="${v%%:*}"; envs=""
  if [ "$v" != "base" ]; then envs="$(echo "${v#*:}" | tr ',' ' ')"; fi
  env $envs timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_$name.json > gpurun_out/bench_$name.log 2>&1
  echo "bench $name [$envs] rc=$? $(tail -1 gpurun_out/bench_$name.log | cut -c1-110)"
done
