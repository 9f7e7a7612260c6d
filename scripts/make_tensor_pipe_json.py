#!/usr/bin/env python
"""profiles/<tag>_ncu_*.csv (extracts of `ncu --set full` captures) -> profiles/tensor_pipe.json:
{tier: {"source": ..., "captures": {class: {kernel, tensor_pipe_active_pct, duration_us, dram_MB}}}},
attached to the bench line as roofline.ncu_tensor_pipe (committed ncu evidence, not a live measurement).

usage: python scripts/make_tensor_pipe_json.py <tag> [<later tag> ...]   (later tags override earlier ones)
"""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tags = sys.argv[1:] or ["r2a"]
DESC = {"c32t": "32x32 256->256 3x3 +temb", "c32cat": "32x32 512->256 3x3 +temb (largest class)",
        "c16": "16x16 256->256 3x3 +res", "c8": "8x8 256->256 3x3 +res", "qkv16": "16x16 256->768 1x1 (q|k|v)",
        "g32": "32x32 256->256 3x3 GroupNorm-on-load +temb", "g32cat": "32x32 512->256 3x3 GroupNorm-on-load +temb",
        "g16": "16x16 256->256 3x3 GroupNorm-on-load +temb",
        "g32sc": "32x32 256->256 3x3 GroupNorm-on-load + fused 1x1 shortcut over 512 raw channels",
        "g16sc": "16x16 256->256 3x3 GroupNorm-on-load + fused 1x1 shortcut over 512 raw channels",
        "g32res": "32x32 256->256 3x3 GroupNorm-on-load +res", "attn_tc": "attention 16x16 (+ fused NIN_3 projection)"}
out = {"bf16x3": {"captures": {}}, "bf16": {"captures": {}}}
paths = [(t, p) for t in tags for p in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{t}_ncu_*.csv")))]
for tag, path in paths:
    name = os.path.basename(path)[len(tag) + 5:-4]
    tier = "bf16x3" if name.startswith(("x3_", "bf16x3_")) else "bf16"
    cls = name.split("_", 1)[1]
    rows = list(csv.reader(open(path)))
    h, r = rows[0], rows[2]
    g = lambda k: r[h.index(k)] if k in h else None
    tp = g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    if tp is None or float(tp) == 0.0:
        continue
    unit = rows[1][h.index("gpu__time_duration.sum")]
    dur = float(g("gpu__time_duration.sum")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(unit, 1.0)
    out[tier]["captures"][cls] = {
        "shape": DESC.get(cls, cls), "kernel": g("Kernel Name").split("(")[0].replace("void ", ""),
        "tensor_pipe_active_pct": round(float(tp), 1), "duration_us": round(dur, 1),
        "file": os.path.basename(path)}
for tier in out:
    out[tier]["source"] = (f"ncu --set full --clock-control none, one launch per class selected BY SHAPE "
                           f"(scripts/one_op.py at B=256; attention from a real step), {'+'.join(tags)}; metric "
                           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
json.dump(out, open(os.path.join(ROOT, "profiles", "tensor_pipe.json"), "w"), indent=1)
for tier, d in out.items():
    for k, v in d["captures"].items():
        print(tier, k, v["tensor_pipe_active_pct"], "%", v["duration_us"], "us")
