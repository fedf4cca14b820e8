#!/usr/bin/env bash
# Profile pass for profiles/: bench lines, per-layer table, ncu launch list (time + DRAM bytes) of the
# same command, full ncu captures of groups of layers (NCU_GROUPS="skip count ..." pairs; default four
# groups); summaries are made on the box so that only small files travel back (gpurun merges <= 64 MiB).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers_auto.json > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_auto.log | cut -c1-200
if [ -z "${SKIP_ARMS}" ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --variant shift --no-cpu-baseline > gpurun_out/bench_shift.log 2>&1; echo "bench shift rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference rc=$?"
fi
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
i=0
set -- ${NCU_GROUPS:-54 6 66 8 80 8 98 6}
while [ $# -ge 2 ]; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s $1 -c $2 -o gpurun_out/prof_g$i -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_g$i.log 2>&1; echo "ncu full group $i rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_g$i.ncu-rep gpurun_out/layers_auto.json $(( $1 - 54 )) > gpurun_out/ncu_full_g$i.md 2>&1
  python tools/ncu_sass_hist.py gpurun_out/prof_g$i.ncu-rep ::regex:conv_mma:2 40 > gpurun_out/sass_hist_g${i}_k2.txt 2>&1
  python tools/ncu_sass_hist.py gpurun_out/prof_g$i.ncu-rep ::regex:conv_mma:5 40 > gpurun_out/sass_hist_g${i}_k5.txt 2>&1
  i=$((i+1)); shift 2
done
rm -f gpurun_out/prof_g1.ncu-rep gpurun_out/prof_g2.ncu-rep gpurun_out/prof_g3.ncu-rep
du -sm gpurun_out
