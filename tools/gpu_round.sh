#!/usr/bin/env bash
# One GPU visit: smoke, GPU parity tests, benches, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-auto}; do
timeout 600 python bench.py --steps 10 --warmup 3 --variant $v --layers-out gpurun_out/layers_$v.json ${BENCH_EXTRA} > gpurun_out/bench_$v.log 2>&1; echo "bench $v rc=$?" | tee -a gpurun_out/bench_$v.log
tail -2 gpurun_out/bench_$v.log
done
if [ -n "${NCU}" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s 54 -c 5 -o gpurun_out/prof_mma -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
