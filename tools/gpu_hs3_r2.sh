#!/usr/bin/env bash
# Round-2: parity of the streamed-weight halo mode, library A/B (VARIANTS = tools/micro/bin/ab_<v>), MMA-thread tick trace
mkdir -p gpurun_out
T=${TAG:-hs3}
if [ -z "${SKIP_TESTS}" ]; then
timeout 1200 python -m pytest tests/test_gpu_mma.py tests/test_gpu_resnet50.py tests/test_vgg16.py -m gpu -q ${PYTEST_EXTRA:--x} > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
fi
for v in ${VARIANTS:-base}; do
  if [ "$v" = base ]; then unset TF2B_LIB; else export TF2B_LIB=$PWD/tools/micro/bin/ab_$v/libtf2b200.so; fi
  for n in ${NETS:-resnet50}; do
    timeout 600 python bench.py --net $n --steps 20 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_${n}_${T}_$v.json > gpurun_out/bench_${n}_${T}_$v.log 2>&1; echo "bench $v $n rc=$?"
    tail -1 gpurun_out/bench_${n}_${T}_$v.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$v $n', round(l['value']), l['ms_per_step'], round(l['e2e']['value']))"
  done
done
if [ -n "${TRACE}" ]; then
  export TF2B_LIB=$PWD/tools/micro/bin/exp_hs/libtf2b200.so
  for hs in 1 0; do
    TF2B_MMA_HSTREAM=$hs TF2B_MMA_DEBUG=1 TF2B_MMA_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --executor 0 --no-cpu-baseline 2>&1 | grep -A2 "mma dbg" | grep -v "^--" | awk '/mma dbg/{k=$3 $4 $5 $6 $7; p=!seen[k]++} p' > gpurun_out/dbg_${T}_hs$hs.txt
  done
fi
