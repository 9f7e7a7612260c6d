"""TEST INFRASTRUCTURE — host interpreter for ``psld_op`` programs.

Executes a *dry* plan (``psld_b200.program.build_plan(..., dry=True)``, host tensors) op by op
with torch-CPU arithmetic, following the op semantics documented in ``include/psld_b200.h``.
It validates the HOST logic (layer order, weight packing, buffer wiring, epilogue flags, temb
offsets) without a GPU.  It is not a product path and nothing in ``psld_b200`` imports it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.psld_oracle import upfirdn2d  # noqa: E402
from psld_b200 import _lib as L  # noqa: E402


def _f(t):
    return t.to(torch.float32)


def run_plan(plan, x, t):
    """x: [B,in_ch,H,W] f32, t: [nt] f32 (tau, or log tau for a ``logged`` plan) -> eps."""
    T = {k.data_ptr(): k for k in plan.keep}
    g = lambda p: None if not p else T[p]
    plan.x_in.copy_(x)
    plan.time_buf.copy_(t)
    for op in plan.ops:
        i, f = op.i, op.f
        if op.kind == L.OP_ZERO:
            if op.out[0]:
                ch = next(c for c in plan.stat_chunks if c.data_ptr() == op.out[0])
                nbytes = i[0] | (i[1] << 31)
                assert nbytes % 8 == 0 and nbytes // 8 <= ch.numel()
                ch[: nbytes // 8].zero_()
        elif op.kind == L.OP_LAYOUT:
            src, dst = g(op.inp[0]), g(op.out[0])
            if i[L.LAYOUT_DIR] == 0:
                dst.zero_()
                dst[..., : i[L.LAYOUT_C]].copy_(src.permute(0, 2, 3, 1))
            else:
                dst.copy_(_f(src).permute(0, 3, 1, 2))
        elif op.kind == L.OP_TEMB:
            nf = i[L.TEMB_NF]
            tt = _f(g(op.inp[0]))
            if i[L.TEMB_EMB] == 0:
                lt = tt if i[L.TEMB_LOGGED] else torch.log(tt)
                xp = lt[:, None] * g(op.inp[1])[None, :] * 2 * np.pi
                emb = torch.cat([torch.sin(xp), torch.cos(xp)], -1)
            else:
                half = nf // 2
                e = torch.exp(torch.arange(half, dtype=torch.float32) * -(np.log(10000.0) / (half - 1)))
                e = tt[:, None] * e[None, :]
                emb = torch.cat([torch.sin(e), torch.cos(e)], 1)
            h = F.linear(emb, g(op.inp[2]), g(op.inp[3]))
            h = F.linear(F.silu(h), g(op.inp[4]), g(op.inp[5]))
            g(op.out[0]).copy_(F.linear(F.silu(h), g(op.inp[6]), g(op.inp[7])))
        elif op.kind == L.OP_GN:
            x1, x2 = g(op.inp[0]), g(op.inp[1])
            xx = _f(x1) if x2 is None else torch.cat([_f(x1), _f(x2)], -1)
            assert xx.shape[-1] == i[L.GN_C1] + i[L.GN_C2]
            y = F.group_norm(xx.permute(0, 3, 1, 2), i[L.GN_G], g(op.inp[2]), g(op.inp[3]), eps=f[0])
            if op.inp[4]:      # producer-side micro-group statistics must describe exactly x1 / x2
                for src, st in ((x1, g(op.inp[4])), (x2, g(op.inp[5]))):
                    if src is None:
                        continue
                    v = _f(src).double().reshape(src.shape[0], -1, src.shape[-1] // 4, 4)
                    ref = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1)
                    assert st is not None and tuple(st.shape) == tuple(ref.shape), (st.shape, ref.shape)
                    # bf16 plans: the producer summed unrounded fp32 values, tolerate the rounding
                    assert float((st - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 1e-3
            if i[L.GN_AFFINE_ONLY]:
                xn = xx.permute(0, 3, 1, 2)
                Bn, Cn = xn.shape[:2]
                G = i[L.GN_G]
                v = xn.reshape(Bn, G, -1).double()
                mean, var = v.mean(-1), v.var(-1, unbiased=False)
                rstd = 1.0 / torch.sqrt(var + f[0])
                cpg = Cn // G
                sc = rstd.repeat_interleave(cpg, 1) * g(op.inp[2]).double()[None]
                sh = g(op.inp[3]).double()[None] - mean.repeat_interleave(cpg, 1) * sc
                g(op.out[0]).copy_(torch.stack([sc, sh], -1).float())
                continue
            if i[L.GN_SILU]:
                y = F.silu(y)
            g(op.out[0]).copy_(y.permute(0, 2, 3, 1))
        elif op.kind == L.OP_AXPBY:
            n = i[0] | (i[1] << 31)
            xa, ya, dst = g(op.inp[0]), g(op.inp[1]), g(op.out[0])
            assert xa.numel() == n == dst.numel() and (ya is None or ya.numel() == n)
            a, b = torch.tensor(f[0], dtype=torch.float32), torch.tensor(f[1], dtype=torch.float32)
            dst.copy_(a * xa if ya is None else a * xa + b * ya)
        elif op.kind == L.OP_FIR:
            KH = i[L.FIR_KH]
            k = np.asarray([f[j] for j in range(KH * KH)], np.float32).reshape(KH, KH)
            xx = _f(g(op.inp[0])).permute(0, 3, 1, 2)
            y = upfirdn2d(xx, k, up=i[L.FIR_UP], down=i[L.FIR_DOWN], pad=(i[L.FIR_PAD0], i[L.FIR_PAD1]))
            g(op.out[0]).copy_(y.permute(0, 2, 3, 1))
        elif op.kind == L.OP_CONV:
            ks, Cin = i[L.CONV_KS], i[L.CONV_C1] + i[L.CONV_C2]
            cout = i[L.CONV_COUT]
            x1, x2 = g(op.inp[0]), g(op.inp[1])
            assert i[L.CONV_IN_LAYOUT] == L.NHWC
            xx = _f(x1) if x2 is None else torch.cat([_f(x1), _f(x2)], -1)
            assert xx.shape[-1] == Cin
            if op.engine == L.ENGINE_TC_GN:      # GroupNorm(+SiLU) applied on load, bf16-rounded
                aff = g(op.inp[6])
                xx = xx * aff[:, None, None, :, 0] + aff[:, None, None, :, 1]
                if i[L.CONV_GN_SILU]:
                    xx = F.silu(xx)
                xx = xx.to(torch.bfloat16).float()
            w = _f(g(op.inp[4]))
            w_ext, x_ext = None, None
            if op.engine in (L.ENGINE_TC, L.ENGINE_TC_GN) and i[L.CONV_EXT_C1] > 0:    # fused 1x1 shortcut (K-extension)
                ne = i[L.CONV_EXT_C1] + i[L.CONV_EXT_C2]
                w_ext = w[:, ks * ks * Cin:].reshape(cout, ne, 1, 1)
                w = w[:, : ks * ks * Cin]
                e1, e2 = g(op.inp[8]), g(op.inp[9])
                x_ext = _f(e1) if e2 is None else torch.cat([_f(e1), _f(e2)], -1)
                assert x_ext.shape[-1] == ne
            if op.engine in (L.ENGINE_TC, L.ENGINE_TC_GN):
                w = w.reshape(cout, ks, ks, Cin).permute(0, 3, 1, 2)
            else:
                w = w.reshape(ks, ks, Cin, cout).permute(3, 2, 0, 1)
            bias = g(op.inp[5])
            y = F.conv2d(xx.permute(0, 3, 1, 2), w, bias, stride=i[L.CONV_STRIDE], padding=i[L.CONV_PAD])
            assert y.shape[2] == i[L.CONV_OH] and y.shape[3] == i[L.CONV_OW]
            if w_ext is not None:
                y = y + F.conv2d(x_ext.permute(0, 3, 1, 2), w_ext)
            if op.inp[3]:
                tp = g(op.inp[3])
                off = i[L.CONV_TEMB_OFF]
                rows = tp[:, off:off + cout]
                if i[L.CONV_TEMB_BSTRIDE] == 0:
                    rows = rows[:1].expand(y.shape[0], -1)
                else:
                    assert i[L.CONV_TEMB_BSTRIDE] == tp.shape[1]
                y = y + rows[:, :, None, None]
            if op.inp[2]:
                y = y + _f(g(op.inp[2])).permute(0, 3, 1, 2)
            y = y * f[0]
            out = g(op.out[0])
            if op.out[1]:      # accumulated (the program zeroes the arena first)
                v = y.permute(0, 2, 3, 1).double().reshape(y.shape[0], -1, cout // 4, 4)
                g(op.out[1]).add_(torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1))
            if i[L.CONV_OUT_LAYOUT] == L.NCHW:
                valid = int(f[1]) if op.engine in (L.ENGINE_TC, L.ENGINE_TC_GN) else cout
                out.copy_(y[:, :valid])
            else:
                out.copy_(y.permute(0, 2, 3, 1))
        elif op.kind == L.OP_ATTN:
            Cc = i[L.ATTN_C]
            qkv = _f(g(op.inp[0]))
            B = qkv.shape[0]
            qkv = qkv.reshape(B, -1, 3 * Cc)
            q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
            w = torch.softmax(torch.einsum("bqc,bkc->bqk", q, k) * f[0], dim=-1)
            o = torch.einsum("bqk,bkc->bqc", w, v)
            out = g(op.out[0])
            if i[L.ATTN_PROJ]:      # fused NIN_3 + skip connection (+ statistics of the result)
                assert op.engine == L.ENGINE_TC and i[L.ATTN_HW] % 128 == 0
                y = F.linear(o.to(torch.bfloat16).float(), _f(g(op.inp[1])), g(op.inp[2]))
                y = (y + _f(g(op.inp[3])).reshape(y.shape)) * f[1]
                if op.out[1]:
                    v4 = y.double().reshape(B, -1, Cc // 4, 4)
                    g(op.out[1]).add_(torch.stack([v4.sum((1, 3)), (v4 * v4).sum((1, 3))], -1))
                o = y
            out.copy_(o.reshape(out.shape))
        else:
            raise AssertionError(f"unknown op kind {op.kind}")
    return plan.eps.clone()
