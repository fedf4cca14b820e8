#!/usr/bin/env bash
# Round-2: ncu of the two bandwidth kernels of the stem (space-to-depth, max-pool) inside a bench step
mkdir -p gpurun_out
T=${TAG:-aux}
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"raw224|maxpool" -c 6 --csv --log-file gpurun_out/aux_$T.csv python bench.py --net ${NET:-resnet50} --steps 1 --warmup 1 --no-cpu-baseline --executor 0 > gpurun_out/aux_$T.log 2>&1; echo "ncu rc=$?"
