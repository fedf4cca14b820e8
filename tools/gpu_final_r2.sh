#!/usr/bin/env bash
# Round-2 final measurement pass (one B200): bench lines of every BASELINE config, the reference arm, the ncu launch
# lists of the same commands (plain launches so that the launch order is the layer order), full ncu captures of
# the dominant kernels, summarised on the box.  Outputs: gpurun_out/final_*.
mkdir -p gpurun_out
export TRAFFIC_DIR=gpurun_out
T=final
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt 2>&1; nproc >> gpurun_out/${T}_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/${T}_layers_resnet50.json > gpurun_out/${T}_bench_resnet50.log 2>&1; echo "bench resnet50 rc=$?"
timeout 900 python bench.py --variant mma --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_resnet50_mma.log 2>&1; echo "bench mma rc=$?"
timeout 900 python bench.py --variant shift --steps 10 --warmup 3 --layers-out gpurun_out/${T}_layers_shift.json > gpurun_out/${T}_bench_shift.log 2>&1; echo "bench shift rc=$?"
timeout 900 python bench.py --net vgg16 --steps 20 --warmup 3 --layers-out gpurun_out/${T}_layers_vgg16.json > gpurun_out/${T}_bench_vgg16.log 2>&1; echo "bench vgg16 rc=$?"
timeout 900 python bench.py --net googlenet --steps 20 --warmup 3 --layers-out gpurun_out/${T}_layers_googlenet.json > gpurun_out/${T}_bench_googlenet.log 2>&1; echo "bench googlenet rc=$?"
timeout 900 python bench.py --net squeezenet --steps 20 --warmup 3 --layers-out gpurun_out/${T}_layers_squeezenet.json > gpurun_out/${T}_bench_squeezenet.log 2>&1; echo "bench squeezenet rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; echo "bench reference rc=$?"
for f in resnet50 resnet50_mma shift vgg16 googlenet squeezenet reference; do tail -1 gpurun_out/${T}_bench_$f.log | cut -c1-180; done
# launch lists
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_resnet50.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_traffic.py gpurun_out/${T}_launches_resnet50.csv 256 resnet50 conv_mma > gpurun_out/${T}_launches_resnet50_summary.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_shift.csv python bench.py --variant shift --steps 2 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/${T}_ncu_list_shift.log 2>&1; echo "ncu list shift rc=$?"
python tools/ncu_traffic.py gpurun_out/${T}_launches_shift.csv 256 resnet50 conv_sa > gpurun_out/${T}_launches_shift_summary.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_vgg16.csv python bench.py --net vgg16 --steps 2 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/${T}_ncu_list_vgg.log 2>&1; echo "ncu list vgg rc=$?"
python tools/ncu_traffic.py gpurun_out/${T}_launches_vgg16.csv 128 vgg16 conv_mma > gpurun_out/${T}_launches_vgg16_summary.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_googlenet.csv python bench.py --net googlenet --steps 2 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/${T}_ncu_list_goog.log 2>&1; echo "ncu list googlenet rc=$?"
python tools/ncu_traffic.py gpurun_out/${T}_launches_googlenet.csv 64 googlenet conv_mma > gpurun_out/${T}_launches_googlenet_summary.txt 2>&1
# full captures: conv_mma launches of the timed step (54 = layer 0): L0-L5, L11-L16, L26-L33, L43-L52
i=0
set -- 54 6 65 6 80 8 97 10
while [ $# -ge 2 ]; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s $1 -c $2 -o gpurun_out/${T}_prof_mma_$i -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/${T}_ncu_mma_$i.log 2>&1; echo "ncu mma group $i rc=$?"
  python tools/ncu_summary.py gpurun_out/${T}_prof_mma_$i.ncu-rep gpurun_out/${T}_layers_resnet50.json $(( $1 - 54 )) > gpurun_out/${T}_ncu_mma_$i.md 2>&1
  python tools/ncu_stalls.py gpurun_out/${T}_prof_mma_$i.ncu-rep > gpurun_out/${T}_ncu_mma_stalls_$i.txt 2>&1
  if [ $i = 0 ]; then python tools/ncu_sass_hist.py gpurun_out/${T}_prof_mma_$i.ncu-rep ::regex:conv_mma:5 40 > gpurun_out/${T}_sass_L4.txt 2>&1; fi
  rm -f gpurun_out/${T}_prof_mma_$i.ncu-rep
  i=$((i+1)); shift 2
done
# kernel A: L3 (3x3 64), L12, L28, L32
i=0
set -- 57 1 66 1 82 1 86 1
while [ $# -ge 2 ]; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_sa -s $1 -c $2 -o gpurun_out/${T}_prof_sa_$i -f python bench.py --variant shift --steps 1 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/${T}_ncu_sa_$i.log 2>&1; echo "ncu sa $i rc=$?"
  python tools/ncu_summary.py gpurun_out/${T}_prof_sa_$i.ncu-rep > gpurun_out/${T}_ncu_sa_$i.md 2>&1
  python tools/ncu_stalls.py gpurun_out/${T}_prof_sa_$i.ncu-rep > gpurun_out/${T}_ncu_sa_stalls_$i.txt 2>&1
  rm -f gpurun_out/${T}_prof_sa_$i.ncu-rep
  i=$((i+1)); shift 2
done
du -sm gpurun_out
