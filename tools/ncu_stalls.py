#!/usr/bin/env python
"""Warp-stall / pipe breakdown of the launches of an `ncu --set full` report: for every launch the issue-slot
utilisation, the warp-state sampling ratios (why eligible warps were not issuing), LSU data-pipe wavefronts by
origin, and the pipe utilisations.  Usage: ncu_stalls.py report.ncu-rep > out.txt"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = raw[raw.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    want = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    want += [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled")]
    extra = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
             "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
             "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tmem.sum",
             "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
             "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
             "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
             "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for k, r in enumerate(body):
        name = r[ix["Kernel Name"]]
        for tag in ("conv_mma_kernel", "conv_sa_kernel"):
            if tag in name:
                name = name[name.find(tag):]
        print(f"== launch {k}: {name.split('(')[0]}  grid {r[ix['Grid Size']] if 'Grid Size' in ix else ''}")
        for m in extra:
            if m in ix:
                print(f"   {m:90s} {r[ix[m]]} {units[ix[m]]}")
        st = []
        for m in want:
            try:
                st.append((float(r[ix[m]].replace(",", "")), m))
            except ValueError:
                pass
        for v, m in sorted(st, reverse=True)[:14]:
            print(f"   {v:10.3f}  {m}")


if __name__ == "__main__":
    main()
