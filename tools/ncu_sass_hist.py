#!/usr/bin/env python
"""Per-opcode executed-instruction histogram and hottest SASS lines of one kernel of an ncu report.
Usage: ncu_sass_hist.py report.ncu-rep <kernel-id filter, e.g. ::regex:conv_mma:1> [top]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
print(lines[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
ops = Counter()
samples = Counter()
tot = 0
body = rows[1:]
for r in body:
    src = r[ix["Source"]].strip()
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("IMAD", "SHF", "LDS", "STS", "LDG", "STG") else "")
    ops[op] += n
    samples[op] += s
    tot += n
print("total warp instructions", tot)
for op, n in ops.most_common(top):
    print(f"{op:16s} {n:12d} {100.0*n/tot:5.1f}%  samples {samples[op]}")
# hottest individual SASS lines by warp-stall samples (where the warps sit), with the stall-reason columns the
# source page carries
stall_cols = [h for h in hdr if h.lower().startswith("stall_") or "stall" in h.lower()]
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(f"total samples {tot_s}; stall columns: {stall_cols[:24]}")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in order:
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    why = []
    for h in stall_cols:
        try:
            v = int(r[ix[h]] or 0)
        except ValueError:
            continue
        if v > 0 and h != "# Samples":
            why.append((v, h))
    why = " ".join(f"{h}={v}" for v, h in sorted(why, reverse=True)[:3])
    print(f"{i:6d} samples {s:7d} ({100.0*s/max(1,tot_s):4.1f}%) exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {why}")
