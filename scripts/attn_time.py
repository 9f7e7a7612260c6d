#!/usr/bin/env python
"""Times the 16x16 attention (+ fused projection) op in isolation (B = 256, L2 flushed): ONE_OP_X3=1 for the split tier."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from psld_b200 import _lib as L  # noqa: E402
import _ops  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()
B, hw, Cc = int(os.environ.get("ONE_OP_B", "256")), 16, 256
x3 = os.environ.get("ONE_OP_X3", "0") == "1"
g = torch.Generator(device="cpu").manual_seed(0)
act = (lambda t: _ops.to_split(t.to(dev))) if x3 else (lambda t: t.to(dev, torch.bfloat16))
qkv = act(torch.randn(B, hw, hw, 3 * Cc, generator=g))
x = act(torch.randn(B, hw, hw, Cc, generator=g))
op, out, keep = _ops.attn_op(qkv, Cc, engine=L.ENGINE_TC, proj=(torch.randn(Cc, Cc, generator=g) * 0.05,
                                                               torch.randn(Cc, generator=g), x, 0.7071))
L.check(lib.psld_op_prepare(op), "prepare")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(23):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.psld_op_run(op, L.stream_ptr()), "run")
    e1.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"attention 16x16 C=256 B={B} {'bf16x3' if x3 else 'bf16'}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")
lib.psld_op_release(op)
