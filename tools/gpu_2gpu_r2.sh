#!/usr/bin/env bash
# Round-2: both bench arms on two GPUs of one box, launched the way the driver launches them
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench $N gpus rc=$?"
grep '^{"metric"' gpurun_out/bench_${N}gpu.log | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_${N}gpu_reference.log 2>&1; echo "reference $N gpus rc=$?"
grep '^{"metric"' gpurun_out/bench_${N}gpu_reference.log | tail -1 | cut -c1-300
