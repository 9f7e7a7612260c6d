#!/usr/bin/env python
"""gpurun_out/traffic.csv (ncu dram bytes + duration per launch of one bench step) -> profiles/traffic.json"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = [l for l in open(os.path.join(ROOT, "gpurun_out", "traffic.csv")) if not l.startswith("==")]
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("psld::", "").split("<")[0]
    v = float(row["Metric Value"].replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
    agg[name][row["Metric Name"]] += v * mult
    if row["Metric Name"] == "gpu__time_duration.sum":
        cnt[name] += 1
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                 "--clock-control none over the timed step of `python bench.py --steps 1 --warmup 3 "
                 "--e2e-nfe 0 --no-cpu-baseline --profile-ops 0 --no-graph` (B=256, bf16), " + (sys.argv[1] if len(sys.argv) > 1 else "round 1"),
       "kernels": {k: {"launches": cnt[k],
                       "dram_read_bytes_per_launch": v["dram__bytes_read.sum"] / cnt[k],
                       "dram_write_bytes_per_launch": v["dram__bytes_write.sum"] / cnt[k],
                       "traffic_bytes_per_launch": (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / cnt[k],
                       "ncu_ms_total": v["gpu__time_duration.sum"] / 1e6} for k, v in agg.items()}}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
for k, v in out["kernels"].items():
    print(k, v["launches"], round(v["traffic_bytes_per_launch"] / 1e6, 2), "MB/launch", round(v["ncu_ms_total"], 3), "ms")
