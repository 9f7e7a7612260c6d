"""Classifier-free guidance as a score_fn: eps = (1 + w) eps_cond - w eps_uncond.

BASELINE configs[4] asks for "class-conditional sampling with classifier-free guidance (two
score_fn passes per step)".  The reference has no such sampler and its NCSN++ takes no label
(``song_sde/ncsnpp.py:288`` "TODO: Add label and other forms of conditioning here!"), so there is
nothing to be bit-compatible with; what IS fixed by the reference is the score_fn call surface
(``score_fn(u, t) -> eps``, ``main/samplers/sde.py:320``, ``main/models/sde/psld.py:354``).  Guidance
is therefore built as a score_fn, not as a sampler: :class:`ClassifierFreeGuidance` wraps two
NCSN++ networks of the same configuration (the conditional one - e.g. the class's own checkpoint -
and the unconditional one) and every sampler of this package (``sscs_sde``, ``em_sde``, ``bb_ode``,
``ip_em_sde``, ``cc_em_sde``) runs it unchanged: its :meth:`plan` returns ONE program

    [ cond program | net_in copy | uncond program | eps = (1 + w) eps_c - w eps_u ]

that ``psld_sampler_run`` replays (and captures into the per-step CUDA graph) exactly like a single
network's program.  Checkable properties: w = 0 reproduces the conditional network's
sampler bit for bit; the combination equals the PyTorch composition of two reference forwards.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L
from .ncsnpp import NCSNpp
from .registry import register_module


def is_native(score_fn) -> bool:
    """Does ``score_fn`` compile to an op program (``plan()``) the native loops can replay?"""
    return isinstance(score_fn, (NCSNpp, ClassifierFreeGuidance))


class GuidedPlan:
    """Concatenation of two compiled programs plus the guidance combination.  The sub-plans own
    their buffers and per-op host state; this object only holds copies of their op records."""

    def __init__(self, a, b, weight: float):
        if (a.B, a.nt, a.logged) != (b.B, b.nt, b.logged) or tuple(a.eps.shape) != tuple(b.eps.shape) \
                or tuple(a.x_in.shape) != tuple(b.x_in.shape):
            raise ValueError("classifier-free guidance: the two networks must share batch, input and output shape")
        self.lib = L.lib()
        self.a, self.b, self.weight = a, b, float(weight)
        self.B, self.nt, self.logged, self.dev, self.dry = a.B, a.nt, a.logged, a.dev, a.dry
        self.bf16, self.x3, self.tc = a.bf16, a.x3, a.tc
        self.x_in, self.time_buf = a.x_in, a.time_buf
        self.eps = torch.empty_like(a.eps)
        self.keep = list(a.keep) + list(b.keep) + [self.eps]
        self.stat_chunks = list(a.stat_chunks) + list(b.stat_chunks)
        ops = [L.Op.from_buffer_copy(a.op_array[k]) for k in range(a.n_ops)]
        ops.append(self._axpby(b.x_in, 1.0, a.x_in, 0.0, None))       # same network input for both
        off = len(ops)
        for k in range(b.n_ops):
            op = L.Op.from_buffer_copy(b.op_array[k])
            if op.kind == L.OP_TEMB:
                op.inp[0] = a.time_buf.data_ptr()                     # one time buffer feeds both
            ops.append(op)
        ops.append(self._axpby(self.eps, 1.0 + self.weight, a.eps, -self.weight, b.eps))
        self.ops = ops
        self.n_ops = len(ops)
        self.op_array = (L.Op * self.n_ops)(*ops)
        self.temb_op = a.temb_op
        self.cin_valid = dict(a.cin_valid)
        self.cin_valid.update({k + off: v for k, v in b.cin_valid.items()})
        self.engine_count = {k: a.engine_count.get(k, 0) + b.engine_count.get(k, 0)
                             for k in set(a.engine_count) | set(b.engine_count)}
        self.launches = self.lib.psld_program_launches(self.op_array, self.n_ops)

    @staticmethod
    def _axpby(out, ca, x, cb, y):
        op = L.Op()
        op.kind, op.engine = L.OP_AXPBY, L.ENGINE_SIMT
        n = out.numel()
        op.i[0], op.i[1] = n & 0x7FFFFFFF, n >> 31
        op.f[0], op.f[1] = ca, cb
        op.inp[0] = x.data_ptr()
        op.inp[1] = y.data_ptr() if y is not None else None
        op.out[0] = out.data_ptr()
        return op

    def run(self, stream=None):
        if self.dry:
            raise RuntimeError("dry plan cannot run")
        s = stream if stream is not None else L.stream_ptr(self.dev)
        L.check(self.lib.psld_program_run(self.op_array, self.n_ops, s), "psld_program_run")

    def release(self):          # per-op host state belongs to the sub-plans
        pass


@register_module(category="score_fn", name="cfg_ncsnpp_b200")
class ClassifierFreeGuidance(nn.Module):
    """``score_fn(u, t) = (1 + w) cond(u, t) - w uncond(u, t)`` over two :class:`NCSNpp` networks.

    ``ClassifierFreeGuidance(config)`` builds both networks from ``config.model.score_fn`` (state-dict
    prefixes ``cond.`` / ``uncond.``; ``config.model.score_fn.guidance_weight`` = w, default 0) - the
    form the reference's registry instantiates, ``cls(config)`` (``main/eval/sample.py:50``);
    ``ClassifierFreeGuidance(cond=net_c, uncond=net_u, weight=w)`` wraps existing modules.
    """

    def __init__(self, config=None, cond=None, uncond=None, weight=None):
        super().__init__()
        if (cond is None or uncond is None) and config is None:
            raise ValueError("ClassifierFreeGuidance needs a config or both networks")
        self.cond = cond if cond is not None else NCSNpp(config)
        self.uncond = uncond if uncond is not None else NCSNpp(config)
        if self.cond is self.uncond:
            raise ValueError("ClassifierFreeGuidance: cond and uncond must be two modules "
                             "(with one network the guided score is the network itself)")
        if weight is None:
            sf = config.model.score_fn if config is not None else {}
            weight = (sf.get("guidance_weight", 0.0) if hasattr(sf, "get") else getattr(sf, "guidance_weight", 0.0))
        self.weight = float(weight)
        for k in ("in_ch", "out_ch", "embedding_type", "image_size"):
            if getattr(self.cond, k, None) != getattr(self.uncond, k, None):
                raise ValueError(f"ClassifierFreeGuidance: the two networks differ in {k}")
            setattr(self, k, getattr(self.cond, k, None))
        self._plans = {}

    def __deepcopy__(self, memo):
        import copy
        return type(self)(cond=copy.deepcopy(self.cond, memo), uncond=copy.deepcopy(self.uncond, memo),
                          weight=self.weight)

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        self._plans = {}
        return r

    @property
    def precision(self):
        return self.cond.precision

    def set_precision(self, precision: str):
        self.cond.set_precision(precision)
        self.uncond.set_precision(precision)
        return self

    def invalidate(self):
        self.cond.invalidate()
        self.uncond.invalidate()
        self._plans = {}

    def plan(self, batch: int, nt: int, logged: bool):
        if self.cond.precision != self.uncond.precision:
            raise ValueError("ClassifierFreeGuidance: the two networks must run in the same precision tier")
        a = self.cond.plan(batch, nt, logged)
        b = self.uncond.plan(batch, nt, logged)
        key = (int(batch), int(nt), bool(logged), self.cond.precision)
        g = self._plans.get(key)
        # the sub-plans are rebuilt when their weights change: a guided plan is valid only for the
        # exact pair of sub-plans (and weight) it was assembled from
        if g is None or g.a is not a or g.b is not b or g.weight != self.weight:
            g = GuidedPlan(a, b, self.weight)
            self._plans[key] = g
        return g

    def forward(self, x, time_cond):
        if not x.is_cuda:
            raise RuntimeError("psld_b200.ClassifierFreeGuidance runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 4 or x.shape[1] != self.in_ch:
            raise ValueError(f"expected [B,{self.in_ch},H,W] input, got {tuple(x.shape)}")
        B = x.shape[0]
        p = self.plan(B, B, False)
        with torch.no_grad():
            p.x_in.copy_(x.to(torch.float32))
            p.time_buf.copy_(time_cond.to(torch.float32).reshape(-1).expand(B))
        p.run()
        return p.eps.clone()
