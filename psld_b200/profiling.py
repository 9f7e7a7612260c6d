"""Per-op device timing and algorithmic-work accounting for a compiled program."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

KIND_NAMES = {L.OP_LAYOUT: "layout", L.OP_TEMB: "temb", L.OP_GN: "groupnorm", L.OP_FIR: "fir",
              L.OP_CONV: "conv", L.OP_ATTN: "attention", L.OP_ZERO: "zero", L.OP_AXPBY: "axpby"}


def op_name(op):
    if op.kind == L.OP_CONV:
        return {L.ENGINE_TC: "conv_tc", L.ENGINE_TC_GN: "conv_tc_gn"}.get(op.engine, "conv_simt")
    if op.kind == L.OP_ATTN:
        return "attn_tc" if op.engine == L.ENGINE_TC else "attn_simt"
    return KIND_NAMES.get(op.kind, "?")


def op_flops(op, cin_valid=None) -> float:
    """Algorithmic FLOPs (2*MAC) of dense-contraction ops; 0 for memory-bound ops.  Zero padding is
    not work: ``cin_valid`` = real input channels of a conv over a channel-padded input, and the
    output head counts its valid output channels only."""
    i = op.i
    if op.kind == L.OP_CONV:
        M = i[L.CONV_N] * i[L.CONV_OH] * i[L.CONV_OW]
        K = i[L.CONV_KS] ** 2 * (cin_valid if cin_valid else i[L.CONV_C1] + i[L.CONV_C2])
        if op.engine in (L.ENGINE_TC, L.ENGINE_TC_GN):
            K += i[L.CONV_EXT_C1] + i[L.CONV_EXT_C2]      # fused 1x1 shortcut
        cout = int(op.f[1]) if (op.engine in (L.ENGINE_TC, L.ENGINE_TC_GN) and op.f[1] >= 1) else i[L.CONV_COUT]
        return 2.0 * M * K * cout
    if op.kind == L.OP_ATTN:
        return 4.0 * i[L.ATTN_N] * i[L.ATTN_HW] ** 2 * i[L.ATTN_C]
    return 0.0


def op_bytes(op, elt: int) -> float:
    """Algorithmic HBM bytes of the memory-bound ops (read once + write once)."""
    i = op.i
    if op.kind == L.OP_GN:      # stats pass reads x, apply pass reads x and writes y
        n = i[L.GN_N] * i[L.GN_HW] * (i[L.GN_C1] + i[L.GN_C2])
        if i[L.GN_AFFINE_ONLY]:
            return (0.0 if op.inp[4] else 1.0) * n * elt
        return (2.0 if op.inp[4] else 3.0) * n * elt
    if op.kind == L.OP_FIR:
        KH, up, down = i[L.FIR_KH], i[L.FIR_UP], i[L.FIR_DOWN]
        OH = (i[L.FIR_H] * up + i[L.FIR_PAD0] + i[L.FIR_PAD1] - KH) // down + 1
        OW = (i[L.FIR_W] * up + i[L.FIR_PAD0] + i[L.FIR_PAD1] - KH) // down + 1
        return float(i[L.FIR_N] * i[L.FIR_C] * (i[L.FIR_H] * i[L.FIR_W] + OH * OW) * elt)
    return 0.0


def profile_plan(plan, iters: int = 3, warmup: int = 1):
    """Times every op of the program with CUDA events on the launching stream.
    Returns {name: {"ms": avg ms per program run, "launch_ms": avg per op, "n": ops,
    "flops": algorithmic FLOPs per run, "bytes": algorithmic bytes per run}}."""
    lib = plan.lib
    stream = L.stream_ptr(plan.dev)
    n = plan.n_ops
    acc = [0.0] * n
    for it in range(warmup + iters):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for k in range(n):
            L.check(lib.psld_op_run(C.byref(plan.op_array[k]), stream), "psld_op_run")
            evs[k + 1].record()
        torch.cuda.synchronize(plan.dev)
        if it >= warmup:
            for k in range(n):
                acc[k] += evs[k].elapsed_time(evs[k + 1])
    out = {}
    elt = 2 if plan.bf16 else 4      # (split bf16 = 4 bytes per element)
    classes = {}
    for k in range(n):
        op = plan.op_array[k]
        if op.kind == L.OP_CONV:
            i = op.i
            key = f"{op_name(op)} {i[L.CONV_OH]}x{i[L.CONV_OW]} {i[L.CONV_C1] + i[L.CONV_C2]}->" \
                  f"{i[L.CONV_COUT]} k{i[L.CONV_KS]}s{i[L.CONV_STRIDE]}" \
                  f"{'+res' if op.inp[2] else ''}{'+temb' if op.inp[3] else ''}" \
                  f"{'+sc' + str(i[L.CONV_EXT_C1] + i[L.CONV_EXT_C2]) if (op.engine in (L.ENGINE_TC, L.ENGINE_TC_GN) and i[L.CONV_EXT_C1]) else ''}"
            c = classes.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0})
            c["ms"] += acc[k] / iters
            c["n"] += 1
            c["flops"] += op_flops(op, plan.cin_valid.get(k))
        d = out.setdefault(op_name(op), {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        d["ms"] += acc[k] / iters
        d["n"] += 1
        d["flops"] += op_flops(op, plan.cin_valid.get(k))
        d["bytes"] += op_bytes(op, elt)
    for d in out.values():
        d["launch_ms"] = d["ms"] / max(d["n"], 1)
    for c in classes.values():
        c["tflops"] = c["flops"] / (c["ms"] * 1e-3) / 1e12 if c["ms"] > 0 else 0.0
    out["_conv_classes"] = classes
    return out
