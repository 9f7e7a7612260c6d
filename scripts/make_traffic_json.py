#!/usr/bin/env python
"""gpurun_out/traffic_<tier>.csv (ncu dram bytes + duration per launch of one bench step, per precision
tier; scripts/gpu_evidence.sh) -> profiles/traffic.json {tier: {source, kernels}} and
profiles/<tag>_launches_<tier>.md (per-kernel launch count / time share of the step).

usage: python scripts/make_traffic_json.py <tag>
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out_all = {}
for tier in ("bf16x3", "bf16", "fp32"):
    src = os.path.join(ROOT, "gpurun_out", f"traffic_{tier}.csv")
    if not os.path.exists(src):
        continue
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("psld::", "").split("<")[0]
        v = float(row["Metric Value"].replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        agg[name][row["Metric Name"]] += v * mult
        if row["Metric Name"] == "gpu__time_duration.sum":
            cnt[name] += 1
    if not cnt:
        continue
    out_all[tier] = {
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                  "--clock-control none over the timed step of `python bench.py --steps 1 --warmup 3 "
                  f"--e2e-nfe 0 --profile-ops 0 --no-graph --precision {tier}` (B=256), {tag}",
        "kernels": {k: {"launches": cnt[k],
                        "dram_read_bytes_per_launch": v["dram__bytes_read.sum"] / cnt[k],
                        "dram_write_bytes_per_launch": v["dram__bytes_write.sum"] / cnt[k],
                        "traffic_bytes_per_launch": (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / cnt[k],
                        "ncu_ms_total": v["gpu__time_duration.sum"] / 1e6} for k, v in agg.items()}}
    tot = sum(v["gpu__time_duration.sum"] for v in agg.values()) / 1e6
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches_{tier}.md"), "w") as f:
        f.write(f"# ncu launch list, one bench step, tier {tier} ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none` over the timed step\n"
                "(cold-cache, serialised launches: compare SHARES with `per_kernel_ms` of the bench line, not absolutes).\n\n"
                f"total {tot:.3f} ms over {sum(cnt.values())} launches\n\n"
                "| kernel | launches | ms | share | DRAM MB / launch |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            ms = v["gpu__time_duration.sum"] / 1e6
            mb = (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / cnt[k] / 1e6
            f.write(f"| `{k}` | {cnt[k]} | {ms:.3f} | {ms / tot:.3f} | {mb:.1f} |\n")
    print(tier, "total", round(tot, 3), "ms,", sum(cnt.values()), "launches")
    for k, v in sorted(out_all[tier]["kernels"].items(), key=lambda kv: -kv[1]["ncu_ms_total"]):
        print("  ", k, v["launches"], round(v["traffic_bytes_per_launch"] / 1e6, 2), "MB/launch", round(v["ncu_ms_total"], 3), "ms")
json.dump(out_all, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
