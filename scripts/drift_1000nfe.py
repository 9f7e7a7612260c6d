#!/usr/bin/env python
"""Precision-tier drift over the FULL 1000-NFE SSCS trajectory of BASELINE configs[1] (CIFAR-10 NCSN++,
init_scale = 1 so that the network is numerically visible), GPU vs GPU: the bf16x3 and bf16 tensor-core
tiers against the fp32 CUDA-core tier on identical weights, prior and (Philox) noise.

    python scripts/drift_1000nfe.py [B] [nfe]  ->  one JSON line (also written to gpurun_out/drift_1000nfe.json)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _net import make_net  # noqa: E402
from psld_b200 import PSLD, SSCSSampler, cifar10_config, time_grid  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nfe = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg = cifar10_config(n_discrete_steps=nfe, batch_size=B, n_samples=B)
cfg.model.score_fn.init_scale = 1.0
ts, n = time_grid(cfg)
sde = PSLD(cfg)
u0 = sde.prior_sampling([B, 3, 32, 32])
out, secs = {}, {}
for tier in ("fp32", "bf16x3", "bf16"):
    net, _ = make_net(cfg, tier)
    S = SSCSSampler(cfg, sde, net)
    S.state_dtype = torch.float64
    t0 = time.perf_counter()
    out[tier] = S.sample(u0.cuda(), ts.cuda(), n, denoise=True, eps=1e-3).double().cpu()
    torch.cuda.synchronize()
    secs[tier] = round(time.perf_counter() - t0, 2)
    del net, S
ref = out["fp32"]
rel = lambda a: float((a - ref).norm() / ref.norm())
mx = lambda a: float((a - ref).abs().max() / ref.abs().max())
line = {"what": f"CIFAR-10 NCSN++ (init_scale=1), SSCS, {nfe} NFE, B={B}, fp64 state, identical Philox noise; "
                "end state vs the fp32 CUDA-core tier",
        "bf16x3": {"rel_l2": rel(out["bf16x3"]), "max_abs_over_max_ref": mx(out["bf16x3"])},
        "bf16": {"rel_l2": rel(out["bf16"]), "max_abs_over_max_ref": mx(out["bf16"])},
        "max_abs_ref": float(ref.abs().max()), "finite": bool(torch.isfinite(ref).all()), "seconds": secs}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(line, open(os.path.join(ROOT, "gpurun_out", "drift_1000nfe.json"), "w"), indent=1)
print(json.dumps(line))
