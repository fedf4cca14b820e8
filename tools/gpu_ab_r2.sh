#!/usr/bin/env bash
# A/B of library builds (tools/micro/bin/ab_<name>/libtf2b200.so, TF2B_LIB override): bench line + per-layer table
# each, plus a parity smoke of the flat / residual cases.  VARIANTS="name ..." ("base" = the product library)
mkdir -p gpurun_out
for v in ${VARIANTS:-base}; do
  if [ "$v" = base ]; then unset TF2B_LIB; else export TF2B_LIB=$PWD/tools/micro/bin/ab_$v/libtf2b200.so; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA} --layers-out gpurun_out/layers_ab_$v.json > gpurun_out/bench_ab_$v.log 2>&1; echo "bench $v rc=$?"
  tail -1 gpurun_out/bench_ab_$v.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$v', round(l['value']), l['ms_per_step'], round(l['e2e']['value']))"
  if [ -n "${PARITY}" ]; then
    timeout 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_resnet50.py -q -x -k "flat or residual or bn256 or plan or auto" > gpurun_out/pytest_ab_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 gpurun_out/pytest_ab_$v.log
  fi
done
