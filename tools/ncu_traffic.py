#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernel from the ncu launch list of `bench.py --steps 2
--warmup 1` (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch).
Writes profiles/r02_ncu_traffic[_<net>].json (read back by bench.py as roofline.traffic) and a per-kernel
summary.  Usage: ncu_traffic.py launches.csv [batch] [net] [kernel: conv_mma|conv_sa] > summary.txt"""
import csv
import json
import os
import sys
from collections import defaultdict

path = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
net = sys.argv[3] if len(sys.argv) > 3 else "resnet50"
dom = sys.argv[4] if len(sys.argv) > 4 else "conv_mma"
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
launch = defaultdict(dict)
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    launch[int(r["ID"])]["name"] = r["Kernel Name"]
    launch[int(r["ID"])][r["Metric Name"]] = v * scale
ids = sorted(launch)
first = "raw224" if any("raw224" in launch[i]["name"] for i in ids) else "chw_to_hwc"
starts = [i for i in ids if first in launch[i]["name"]]
step = [i for i in ids if starts[1] <= i < (starts[2] if len(starts) > 2 else ids[-1] + 1)]   # second step = timed step
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in step:
    L = launch[i]
    k = "conv_mma" if "conv_mma" in L["name"] else ("conv_sa" if "conv_sa" in L["name"] else L["name"].split("(")[0].split("::")[-1])
    a = agg[k]
    a[0] += 1
    a[1] += L.get("gpu__time_duration.sum", 0.0)
    a[2] += L.get("dram__bytes_read.sum", 0.0)
    a[3] += L.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none: "
      f"python bench.py --net {net} --steps 2 --warmup 1 --no-cpu-baseline --executor 0; second step (serialised, cold-cache "
      "per-launch times)")
for k in sorted(agg, key=lambda k: -agg[k][1]):
    n, t, rd, wr = agg[k]
    print(f"{k:26s} launches {n:3d}  {t:9.1f} us  share {100*t/tot:5.1f}%  dram read {rd/1e6:9.1f} MB  write {wr/1e6:9.1f} MB  "
          f"({(rd+wr)/max(t,1e-9)/1e3:7.1f} GB/s)")
print(f"total {tot:.1f} us")
n, t, rd, wr = agg[dom]
out = {"kernel": dom, "net": net, "batch": batch, "launches": n, "dram_bytes_per_launch": (rd + wr) / n,
       "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr, "share_of_step_serialised": t / tot,
       "source": os.path.basename(path)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = "r02_ncu_traffic.json" if (net == "resnet50" and dom == "conv_mma") else f"r02_ncu_traffic_{net}{'_shift' if dom == 'conv_sa' else ''}.json"
with open(os.path.join(os.environ.get("TRAFFIC_DIR", os.path.join(root, "profiles")), name), "w") as f:
    json.dump(out, f, indent=1)
