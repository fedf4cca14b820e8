#!/usr/bin/env bash
# Round-2 GPU visit: smoke, GPU parity tests, bench lines for the BASELINE configs, micro probes.
# Env: SKIP_TESTS=1, NETS="resnet50 vgg16 googlenet", MICROS="...", NCU=1, TAG=suffix
mkdir -p gpurun_out
T=${TAG:-a}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_$T.txt 2>&1
nproc >> gpurun_out/gpu_$T.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_$T.log
if [ -z "${SKIP_TESTS}" ]; then
  timeout 2400 python -m pytest tests -m gpu -q ${PYTEST_EXTRA:--x} > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$T.log
  tail -15 gpurun_out/pytest_gpu_$T.log
fi
for n in ${NETS:-resnet50}; do
  timeout 900 python bench.py --net $n --steps 20 --warmup 3 --layers-out gpurun_out/layers_${n}_$T.json ${BENCH_EXTRA} > gpurun_out/bench_${n}_$T.log 2>&1; echo "bench $n rc=$?"
  tail -1 gpurun_out/bench_${n}_$T.log | cut -c1-400
done
if [ -n "${SHIFT}" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --variant shift --no-cpu-baseline --layers-out gpurun_out/layers_shift_$T.json > gpurun_out/bench_shift_$T.log 2>&1; echo "bench shift rc=$?"
  tail -1 gpurun_out/bench_shift_$T.log | cut -c1-300
fi
if [ -n "${REFERENCE}" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$T.log 2>&1; echo "bench reference rc=$?"
  tail -1 gpurun_out/bench_reference_$T.log | cut -c1-300
fi
if [ -n "${NCU}" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_$T.log 2>&1; echo "ncu list rc=$?"
fi
if [ -n "${MICROS}" ]; then
  NCU=${NCU_MICRO} MICROS="${MICROS}" bash tools/gpu_micro.sh
fi
