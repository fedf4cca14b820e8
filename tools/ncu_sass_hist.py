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
