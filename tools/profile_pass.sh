#!/usr/bin/env bash
# Profiling pass for profiles/: bench line (with CPU baseline), per-layer table, ncu launch list of
# the same command, one full ncu capture of the dominant kernel (three representative layers).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers_auto.json > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_final.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --variant shift --no-cpu-baseline > gpurun_out/bench_final_shift.log 2>&1; echo "bench shift rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
# layers 4 (1x1 64->256 + residual), 16 (3x3 128->128), 32 (3x3 256->256, BN=256) of the second step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s 58 -c 29 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
