#!/usr/bin/env bash
# One GPU visit: smoke, GPU parity tests, bench (+CPU baseline, per-layer table), ncu launch list of
# the same command, one full ncu capture of the dominant kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
if [ -z "${SKIP_TESTS}" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers_auto.json ${BENCH_EXTRA} > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_auto.log | cut -c1-600
if [ -n "${NCU}" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s ${NCU_SKIP:-58} -c ${NCU_COUNT:-29} -o gpurun_out/prof_mma -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
