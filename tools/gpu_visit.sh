#!/usr/bin/env bash
# One GPU visit: smoke, GPU parity tests, bench (+CPU baseline, per-layer table), ncu launch list of
# the same command (with DRAM bytes), one full ncu capture of the dominant kernel.  Outputs: gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
if [ -z "${SKIP_TESTS}" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers_auto.json ${BENCH_EXTRA} > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_auto.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --variant shift --no-cpu-baseline > gpurun_out/bench_shift.log 2>&1; echo "bench shift rc=$?"
tail -1 gpurun_out/bench_shift.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference rc=$?"
tail -1 gpurun_out/bench_reference.log | cut -c1-300
if [ -n "${NCU}" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
# one whole step of conv_mma launches (54) after the warm-up step
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s 54 -c 54 -o gpurun_out/prof_mma -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
if [ -n "${MICRO}" ]; then
NCU=${NCU} bash tools/gpu_micro.sh
fi
