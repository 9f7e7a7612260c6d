#!/usr/bin/env python
"""Per-CTA phase timeline of conv_tc_kernel (library built with PSLD_NVCC_EXTRA=-DPSLD_TC_TRACE).
    PSLD_NVCC_EXTRA=-DPSLD_TC_TRACE python -m psld_b200.build --force; python scripts/tc_trace.py c8 c16
    (ONE_OP_X3=1: split-bf16 operands)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from psld_b200 import _lib as L  # noqa: E402
import _ops  # noqa: E402
import one_op  # noqa: E402

NAMES = ["entry", "prologue done", "first stage full (MMA)", "last MMA issued", "last tile: acc ready",
         "last tile: epilogue done", "after final sync"]


def main():
    dev = torch.device("cuda:0")
    lib = L.lib()
    g = torch.Generator(device="cpu").manual_seed(0)
    for name in sys.argv[1:]:
        hw, c1, c2, cout, ks, res, temb, stats, gn = one_op.SHAPES[name]
        B = one_op.B
        act = (lambda t: _ops.to_split(t.to(dev))) if one_op.X3 else (lambda t: t.to(dev, torch.bfloat16))
        x1 = act(torch.randn(B, hw, hw, c1, generator=g))
        x2 = act(torch.randn(B, hw, hw, c2, generator=g)) if c2 else None
        w = torch.randn(cout, c1 + c2, ks, ks, generator=g) * 0.05
        r = act(torch.randn(B, hw, hw, cout, generator=g)) if res else None
        t = torch.randn(B, cout, generator=g).to(dev) if temb else None
        op, out, keep = _ops.conv_op(x1, x2, w, torch.randn(cout, generator=g), residual=r, temb=t,
                                     temb_bstride=cout if temb else 0, engine=L.ENGINE_TC, mg_stats=stats)
        L.check(lib.psld_op_prepare(op), "prepare")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(4):
            flush.zero_()
            e0.record()
            L.check(lib.psld_op_run(op, L.stream_ptr()), "run")
            e1.record()
        torch.cuda.synchronize()
        buf = np.zeros(160 * 8, dtype=np.int64)
        rc = lib.psld_debug_tc_trace(buf.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        tr = buf.reshape(160, 8)[:148]
        mhz = 1.0
        print(f"== {name}: event time {e0.elapsed_time(e1) * 1e3:.1f} us (cycles below; ~1.8 cycles/ns)")
        for cta in (0, 1, 2, 3, 146, 147):
            row = tr[cta]
            print(f" cta {cta:3d}:", " ".join(f"{int(row[k] - row[0]):7d}" if row[k] else "      -" for k in range(7)))
        work = tr[(tr[:, 2] > 0) & (tr[:, 5] > 0)]
        d = work - work[:, :1]
        print(" mean over CTAs with work:")
        for k in range(1, 7):
            print(f"   {NAMES[k]:28s} {d[:, k].mean():9.0f}  (min {d[:, k].min()}, max {d[:, k].max()})")
        L.lib().psld_op_release(op)


if __name__ == "__main__":
    main()
