#!/usr/bin/env python
"""Per-CTA phase timeline of attn_tc_kernel (library built with PSLD_NVCC_EXTRA=-DPSLD_TC_TRACE).
    ONE_OP_X3=1 python scripts/attn_trace.py       # 16x16 attention + fused projection at B = 256"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from psld_b200 import _lib as L  # noqa: E402
import _ops  # noqa: E402

NAMES = ["entry", "Q landed (first MMA)", "S complete", "P written", "O complete", "O normalised",
         "Y complete", "epilogue done"]


def main():
    dev = torch.device("cuda:0")
    lib = L.lib()
    B, hw, Cc = int(os.environ.get("ONE_OP_B", "256")), 16, 256
    x3 = os.environ.get("ONE_OP_X3", "0") == "1"
    g = torch.Generator(device="cpu").manual_seed(0)
    act = (lambda t: _ops.to_split(t.to(dev))) if x3 else (lambda t: t.to(dev, torch.bfloat16))
    qkv = act(torch.randn(B, hw, hw, 3 * Cc, generator=g))
    x = act(torch.randn(B, hw, hw, Cc, generator=g))
    w3 = torch.randn(Cc, Cc, generator=g) * 0.05
    op, out, keep = _ops.attn_op(qkv, Cc, engine=L.ENGINE_TC, proj=(w3, torch.randn(Cc, generator=g), x, 0.7071))
    if os.environ.get("ATTN_NO_STATS", "0") == "1":
        op.out[1] = None
    if os.environ.get("ATTN_NO_RES", "0") == "1":     # only for timing experiments: the kernel then skips the skip connection
        op.inp[3] = None
    L.check(lib.psld_op_prepare(op), "prepare")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(4):
        flush.zero_()
        e0.record()
        L.check(lib.psld_op_run(op, L.stream_ptr()), "run")
        e1.record()
    torch.cuda.synchronize()
    buf = np.zeros(1024 * 8, dtype=np.int64)
    assert lib.psld_debug_attn_trace(buf.ctypes.data_as(C.c_void_p)) == 0
    tr = buf.reshape(1024, 8)[: 2 * B]
    tr = tr[tr[:, 7] > 0]
    d = tr - tr[:, :1]
    print(f"== attention {hw}x{hw} C={Cc} B={B} {'bf16x3' if x3 else 'bf16'}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, "
          f"{len(tr)} CTAs traced (cycles from CTA entry, ~1.9 cycles/ns)")
    for k in range(1, 8):
        print(f"   {NAMES[k]:24s} {d[:, k].mean():9.0f}  (min {d[:, k].min()}, max {d[:, k].max()})")
    lib.psld_op_release(op)


if __name__ == "__main__":
    main()
