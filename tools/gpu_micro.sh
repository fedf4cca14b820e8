#!/usr/bin/env bash
# Runs stand-alone microbenchmarks of tools/micro on the GPU box: MICROS="epi_store param_const" (default).
# Binaries are built here if tools/micro/bin/<name> did not travel.  With NCU=1 each one is also captured with
# the LSU / shared-memory / constant-cache counters the epilogue questions are about.  Outputs: gpurun_out/.
mkdir -p gpurun_out tools/micro/bin
for m in ${MICROS:-epi_store param_const}; do
  if [ ! -x tools/micro/bin/$m ]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/micro/$m.cu -o tools/micro/bin/$m -lcuda \
      > gpurun_out/micro_${m}_build.log 2>&1 || { echo "$m: build failed"; tail -5 gpurun_out/micro_${m}_build.log; continue; }
  fi
  timeout 120 tools/micro/bin/$m > gpurun_out/micro_$m.txt 2>&1; echo "$m rc=$?"
  cat gpurun_out/micro_$m.txt
  if [ -n "${NCU}" ]; then
    timeout 600 ncu --clock-control none --csv --log-file gpurun_out/micro_${m}_ncu.csv --metrics \
gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,\
l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,\
idc__requests.sum,idc__requests_lookup_miss.sum,lts__t_sectors_op_write.sum \
      tools/micro/bin/$m > /dev/null 2>&1; echo "$m ncu rc=$?"
  fi
done
