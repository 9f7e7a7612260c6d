#!/usr/bin/env python
"""Times ONE conv op of the CIFAR-10 plan in isolation (CUDA events, L2 flushed between launches).

    python scripts/one_op.py qkv16 proj16 c8 ...          # named shapes below
Used for kernel experiments and as the ncu target for single-kernel captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from psld_b200 import _lib as L  # noqa: E402
import _ops  # noqa: E402

B = int(os.environ.get("ONE_OP_B", "256"))
X3 = os.environ.get("ONE_OP_X3", "0") == "1"       # split-bf16 operands (bf16x3 tier)
EXT = {"g32sc": 512, "g16sc": 512, "c32sc": 512}    # + fused 1x1 shortcut over a raw input of this many channels (two maps)
SHAPES = {   # name: (HW, C1, C2, Cout, ks, residual, temb, stats, gn)
    "qkv16": (16, 256, 0, 768, 1, False, False, False, False),
    "proj16": (16, 256, 0, 256, 1, True, False, True, False),
    "c32": (32, 256, 0, 256, 3, True, False, True, False),
    "c32t": (32, 256, 0, 256, 3, False, True, True, False),
    "c32cat": (32, 256, 256, 256, 3, False, True, True, False),
    "c16": (16, 256, 0, 256, 3, True, False, True, False),
    "c8": (8, 256, 0, 256, 3, True, False, True, False),
    "c8cat": (8, 256, 256, 256, 3, False, True, True, False),
    "g32": (32, 256, 0, 256, 3, False, True, True, True),
    "g32cat": (32, 256, 256, 256, 3, False, True, True, True),
    "g16": (16, 256, 0, 256, 3, False, True, True, True),
    "g32res": (32, 256, 0, 256, 3, True, False, True, True),
    "g16res": (16, 256, 0, 256, 3, True, False, True, True),
    "g32sc": (32, 256, 0, 256, 3, False, False, True, True),     # GroupNorm_1 -> Conv_1 + Conv_2 shortcut (cat 256+256)
    "g16sc": (16, 256, 0, 256, 3, False, False, True, True),
    "c32sc": (32, 256, 0, 256, 3, False, False, True, False),    # same, unfused conv_tc form
}


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    reps = int(os.environ.get("ONE_OP_REPS", "20"))
    for name in sys.argv[1:]:
        hw, c1, c2, cout, ks, res, temb, stats, gn = SHAPES[name]
        act = (lambda t: _ops.to_split(t.to(dev))) if X3 else (lambda t: t.to(dev, torch.bfloat16))
        x1 = act(torch.randn(B, hw, hw, c1, generator=g))
        x2 = act(torch.randn(B, hw, hw, c2, generator=g)) if c2 else None
        w = torch.randn(cout, c1 + c2, ks, ks, generator=g) * 0.05
        bias = torch.randn(cout, generator=g)
        r = act(torch.randn(B, hw, hw, cout, generator=g)) if res else None
        t = torch.randn(B, cout, generator=g).to(dev) if temb else None
        aff = torch.randn(B, c1 + c2, 2, generator=g).to(dev) if gn else None
        ext = None
        if name in EXT:
            e = EXT[name] // 2
            ext = (act(torch.randn(B, hw, hw, e, generator=g)), act(torch.randn(B, hw, hw, e, generator=g)),
                   torch.randn(cout, 2 * e, 1, 1, generator=g) * 0.05)
        op, out, keep = _ops.conv_op(x1, x2, w, bias, residual=r, temb=t, temb_bstride=cout if temb else 0,
                                     engine=L.ENGINE_TC_GN if gn else L.ENGINE_TC, mg_stats=stats,
                                     affine=aff, ext=ext)
        L.check(L.lib().psld_op_prepare(op), "prepare")
        st = L.stream_ptr()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps + 3):
            flush.zero_()
            if i >= 3:
                ev[i - 3][0].record()
            L.check(L.lib().psld_op_run(op, st), "run")
            if i >= 3:
                ev[i - 3][1].record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
        flops = 2.0 * B * hw * hw * ((c1 + c2) * ks * ks + EXT.get(name, 0)) * cout
        med = ts[len(ts) // 2]
        print(f"{name}: median {med:.1f} us  min {ts[0]:.1f} us  {flops / med * 1e-6:.0f} TFLOP/s  "
              f"finite={bool(torch.isfinite(_ops.val(out).float()).all())}{' (x3: 3 MMAs per product)' if X3 else ''}", flush=True)
        L.lib().psld_op_release(op)


if __name__ == "__main__":
    main()
