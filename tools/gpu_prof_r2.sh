#!/usr/bin/env bash
# Round-2 profile pass: full ncu captures of chosen launches of the timed step (summarised on the box).
# Env: GROUPS_MMA="skip count ..." (launch index among conv_mma launches; 54 = first layer of the 2nd step),
#      GROUPS_SA="skip count ..." (same for conv_sa with --variant shift), TAG
mkdir -p gpurun_out
T=${TAG:-p}
i=0
set -- ${GROUPS_MMA:-54 12}
while [ $# -ge 2 ]; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mma -s $1 -c $2 -o gpurun_out/prof_mma_${T}_$i -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mma_${T}_$i.log 2>&1; echo "ncu mma group $i rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_mma_${T}_$i.ncu-rep > gpurun_out/ncu_mma_${T}_$i.md 2>&1
  python tools/ncu_stalls.py gpurun_out/prof_mma_${T}_$i.ncu-rep > gpurun_out/ncu_mma_stalls_${T}_$i.txt 2>&1
  [ -n "${KEEP_REP}" ] || rm -f gpurun_out/prof_mma_${T}_$i.ncu-rep
  i=$((i+1)); shift 2
done
i=0
set -- ${GROUPS_SA}
while [ $# -ge 2 ]; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_sa -s $1 -c $2 -o gpurun_out/prof_sa_${T}_$i -f python bench.py --variant shift --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sa_${T}_$i.log 2>&1; echo "ncu sa group $i rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_sa_${T}_$i.ncu-rep > gpurun_out/ncu_sa_${T}_$i.md 2>&1
  python tools/ncu_stalls.py gpurun_out/prof_sa_${T}_$i.ncu-rep > gpurun_out/ncu_sa_stalls_${T}_$i.txt 2>&1
  [ -n "${KEEP_REP}" ] || rm -f gpurun_out/prof_sa_${T}_$i.ncu-rep
  i=$((i+1)); shift 2
done
du -sm gpurun_out
