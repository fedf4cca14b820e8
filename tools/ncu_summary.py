#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read with `ncu -i X.ncu-rep --page raw --csv`) into the compact
per-launch table kept under profiles/.  Usage: ncu_summary.py report.ncu-rep [layers.json] > out.md"""
import csv
import io
import json
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1/smem_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_%"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_rt_%"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dsmem_KB"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = raw[raw.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    layers = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in COLS if m in ix]
    print("| # | kernel | layer | " + " | ".join(n for _, n in cols) + " |")
    print("|" + "---|" * (3 + len(cols)))
    for k, r in enumerate(body):
        name = r[ix["Kernel Name"]]
        name = name[name.find("conv_mma_kernel"):] if "conv_mma_kernel" in name else name[:40]
        name = name.split("(")[0]
        lay = ""
        if layers is not None and first + k < len(layers):
            L = layers[first + k]
            lay = f"L{L['layer']} {L['C']}->{L['N']} k{L['k']} s{L['stride']} @{L['OH']}"
        vals = []
        for m, n in cols:
            v = r[ix[m]].replace(",", "")
            u = units[ix[m]]
            try:
                f = float(v)
                if n == "dsmem_KB":
                    f = f / 1024 if u == "byte" else f
                if u == "ns":
                    f /= 1e3
                if u == "byte" and n.endswith("MB"):
                    f /= 1e6
                if u == "Kbyte" and n.endswith("MB"):
                    f /= 1e3
                if u == "Gbyte" and n.endswith("MB"):
                    f *= 1e3
                vals.append(f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}")
            except ValueError:
                vals.append(v)
        print(f"| {k} | {name} | {lay} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
