#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

usage: python scripts/summarize_profiles.py <tag>      (e.g. r1a)
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]


def launches(tag):
    src = os.path.join(G, "launches.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        agg[name][0] += 1
        agg[name][1] += ns
        tot += ns
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, one bench step ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` over the timed step of\n"
                "`python bench.py --steps 1 --warmup 3 --e2e-nfe 0 --no-cpu-baseline --profile-ops 0`\n"
                "(cold-cache, serialised launches: compare SHARES, not absolutes).\n\n"
                f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n\n"
                "| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / tot:.3f} |\n")
    print("wrote", f"{tag}_launches.md")


def full(tag):
    for rep in sorted(glob.glob(os.path.join(G, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        idx = [hdr.index(k) for k in KEYS if k in hdr]
        with open(os.path.join(OUT, f"{tag}_ncu_{name}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            for r in rows:
                w.writerow([r[i] for i in idx])
        print("wrote", f"{tag}_ncu_{name}.csv")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    full(tag)
