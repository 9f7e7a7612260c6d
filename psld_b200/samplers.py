"""SSCS and Euler-Maruyama samplers on the fused sm_100a phase-space kernels.

Drop-in for the reference's ``SSCSSampler`` / ``EulerMaruyamaSampler``
(``main/samplers/sde.py:9-58,227-370``; base class ``main/samplers/base.py:4-31``):

  ``cls(config, sde, score_fn, corrector_fn=None)``
  ``.sample(batch[B,2C,H,W], ts f64[n+1], n_discrete_steps, denoise=True, eps=1e-3) -> Tensor``
  attributes ``.nfe``, property ``.n_steps``; identity corrector by default.

Two execution paths, both on the GPU, neither with a CPU/eager fallback for the update:
  * ``score_fn`` is a :class:`psld_b200.ncsnpp.NCSNpp` (or two of them under
    :class:`psld_b200.guidance.ClassifierFreeGuidance`)  ->  the whole loop runs natively in
    ``psld_sampler_run`` (one C call; per step = network program + ONE fused update kernel);
  * any other callable ``score_fn(u f32, t f32[B]) -> eps``  ->  a thin Python loop that calls
    ``score_fn`` and the same fused kernels (``psld_sscs_update`` / ``psld_em_update``).

Noise: ``sampler.noise`` may hold pre-drawn N(0,1) tensors ``[draws,B,2C,H,W]`` (fp32) in the
reference's ``randn_like`` draw order for parity runs; otherwise the kernels
draw from Philox4x32-10; the key is a function of (``config.evaluation.seed``, global rank at call
time, index of the ``sample()`` call on this sampler): every call gets fresh noise, like the
reference's advancing torch generator seeded with seed + rank (wrapper.py:93-99).
"""
from __future__ import annotations

import abc
import ctypes as C
import os

import torch

from . import _lib as L
from .distributed import call_seed, current_rank
from .guidance import is_native
from .registry import register_module
from .schedule import InpaintTables, PSLDSchedule, StepTables, VPSchedule, VPStepTables


class Sampler(abc.ABC):
    """The abstract sampler (reference base.py:4-31)."""

    def __init__(self, config, sde, score_fn, corrector_fn=None):
        super().__init__()
        self.config = config
        self.sde = sde
        self.score_fn = score_fn
        self.corrector_fn = corrector_fn

    @property
    def n_steps(self):
        return self.config.evaluation.n_discrete_steps

    @abc.abstractmethod
    def predictor_update_fn(self):
        raise NotImplementedError

    def corrector_update_fn(self, x, t, dt):
        if self.corrector_fn is not None:
            return self.corrector_fn(x, t, dt)
        return x, x

    @abc.abstractmethod
    def sample(self):
        raise NotImplementedError


def _opt(config, key, default):
    s = getattr(config.evaluation, "sampler", None)
    try:
        v = s.get(key, None) if hasattr(s, "get") else getattr(s, key, None)
    except Exception:
        v = None
    return default if v is None else v


class _FusedSampler(Sampler):
    KIND = ""

    def __init__(self, config, sde, score_fn, corrector_fn=None):
        super().__init__(config, sde, score_fn, corrector_fn=corrector_fn)
        # the reference's em_sde also drives the VP-SDE baseline (duck-typed on sde.reverse_sde)
        self.vp = None
        if isinstance(sde, VPSchedule) or str(getattr(sde, "type", "")) == "vpsde":
            if self.KIND != "em_sde":
                raise ValueError("the VP-SDE is sampled with em_sde (SSCS is specific to PSLD)")
            self.vp = sde if isinstance(sde, VPSchedule) else VPSchedule.from_sde(sde)
            self.schedule = None
        else:
            self.schedule = sde if isinstance(sde, PSLDSchedule) else PSLDSchedule.from_sde(sde)
        # options (extra keys under evaluation.sampler.*, all optional)
        sd = str(_opt(config, "state_dtype", os.environ.get("PSLD_B200_STATE", "float64")))
        self.state_dtype = torch.float64 if sd in ("float64", "f64", "fp64") else torch.float32
        self.fuse_halves = bool(_opt(config, "fuse_halves", True))
        self.merge_noise = bool(_opt(config, "merge_noise", True))   # Philox mode only
        self.use_graph = bool(_opt(config, "cuda_graph", True))       # replay one captured step
        self.base_seed = int(getattr(config.evaluation, "seed", 0))
        self.calls = 0             # sample() calls so far: folded into the Philox key
        self.seed = call_seed(self.base_seed, current_rank(), 0)
        self.noise = None          # optional pre-drawn noise bank (parity mode)
        self.record = None         # optional [n, B,2C,H,W] buffer filled with per-step states
        self.nfe = 0

    # ------------------------------------------------------------------ reference-named single steps
    # ``sample()`` never calls these (it runs the fused native loop); they exist so that code written
    # against the reference's sampler API (custom loops, correctors, notebooks) keeps working, and
    # they run on the same fused kernels.
    def _scalar(self, v):
        return float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v)

    def _single(self, u, t, dt):
        if self.vp is not None:
            raise NotImplementedError("single-step API is implemented for the PSLD SDE")
        if not u.is_cuda:
            raise RuntimeError("psld_b200 samplers run on CUDA only; there is no CPU path")
        if u.dim() != 4 or u.shape[1] % 2 or (u.shape[1] // 2 * u.shape[2] * u.shape[3]) % 4:
            raise ValueError(f"expected a [B,2C,H,W] phase-space batch, got {tuple(u.shape)}")
        t, dt = self._scalar(t), self._scalar(dt)
        tabs = StepTables(self.schedule, torch.tensor([t], dtype=torch.float64), 1, self.KIND, False,
                          1e-3, self._embedding(), dt=torch.tensor([dt], dtype=torch.float64))
        sdt = torch.float64 if u.dtype == torch.float64 else torch.float32
        state = u.to(sdt).contiguous().clone()
        net_in = state.to(torch.float32)
        B = u.shape[0]
        chw = u.shape[1] // 2 * u.shape[2] * u.shape[3]
        self._single_calls = getattr(self, "_single_calls", 0) + 1
        return tabs, state, net_in, B, chw, L.dtype_code(sdt), L.stream_ptr(u.device)

    def _score(self, net_in, tabs, B):
        with torch.no_grad():
            return self.score_fn(net_in, tabs.tau32[0].to(net_in.device).expand(B)).to(torch.float32).contiguous()

    def predictor_update_fn(self, u, t, dt, z=None):
        """One predictor step from time ``t`` with step ``dt`` (reference sde.py:16-26 for ``em_sde``:
        returns ``(x, x_mean)``; sde.py:331-336 for ``sscs_sde``: returns ``u``).  ``z`` optionally
        supplies the pre-drawn N(0,1) noise (``em_sde``: one ``[B,2C,H,W]`` tensor; ``sscs_sde``: a pair,
        first and second half-step); otherwise the in-kernel Philox generator draws it."""
        lib = L.lib()
        with torch.cuda.device(u.device):
            tabs, state, net_in, B, chw, sdt, stream = self._single(u, t, dt)
            step = (1 << 40) + self._single_calls          # Philox stream ids disjoint from sample()'s
            f32 = lambda v: None if v is None else v.to(u.device, torch.float32).contiguous()
            if self.KIND == "sscs_sde":
                za, zb = (f32(z[0]), f32(z[1])) if z is not None else (None, None)
                L.check(lib.psld_sscs_update(L.ptr(state), L.ptr(state), sdt, L.ptr(net_in), None, L.ptr(za),
                                             None, None, C.byref(tabs.sscs[0]), L.STAGE_HALF_A, self.seed,
                                             step, B, chw, stream), "psld_sscs_update")
                e = self._score(net_in, tabs, B)
                L.check(lib.psld_sscs_update(L.ptr(state), L.ptr(state), sdt, None, L.ptr(e), None,
                                             L.ptr(zb), None, C.byref(tabs.sscs[0]),
                                             L.STAGE_SCORE | L.STAGE_HALF_B, self.seed, step, B, chw,
                                             stream), "psld_sscs_update")
                self._keep = (za, zb, e, tabs)
                return state
            e = self._score(net_in, tabs, B)
            zz = f32(z)
            mean = torch.empty_like(state)
            L.check(lib.psld_em_update(L.ptr(mean), L.ptr(state), sdt, None, L.ptr(e), None, 0,
                                       C.byref(tabs.em[0]), self.seed, step, B, chw, stream),
                    "psld_em_update")
            L.check(lib.psld_em_update(L.ptr(state), L.ptr(state), sdt, None, L.ptr(e), L.ptr(zz),
                                       0 if zz is not None else 1, C.byref(tabs.em[0]), self.seed, step,
                                       B, chw, stream), "psld_em_update")
            self._keep = (zz, e, tabs)
            return state, mean

    def denoising_fn(self, x, t, dt):
        """``x + fbar(x, t) * dt`` without noise (reference sde.py:28-36, 338-348; ``sample()`` calls
        it with ``t = T - eps, dt = eps``)."""
        lib = L.lib()
        with torch.cuda.device(x.device):
            kind, self.KIND = self.KIND, "em_sde"
            try:
                tabs, state, net_in, B, chw, sdt, stream = self._single(x, t, dt)
            finally:
                self.KIND = kind
            e = self._score(net_in, tabs, B)
            L.check(lib.psld_em_update(L.ptr(state), L.ptr(state), sdt, None, L.ptr(e), None, 0,
                                       C.byref(tabs.em[0]), self.seed, 0, B, chw, stream), "psld_em_update")
            self._keep = (e, tabs)
            return state

    def _embedding(self):
        return getattr(self.score_fn, "embedding_type", "fourier")

    def _next_seed(self):
        """Philox key of THIS sample() call (rank resolved now, call counter advanced)."""
        self.seed = call_seed(self.base_seed, current_rank(), self.calls)
        self.calls += 1
        return self.seed

    def _sample_vp(self, batch, ts, n, denoise, eps):
        """Euler-Maruyama on the VP-SDE (state [B,C,H,W]): score_fn + one fused update per step."""
        lib = L.lib()
        if is_native(self.score_fn):
            dev = next(self.score_fn.parameters()).device
        else:
            dev = batch.device if batch.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("psld_b200 samplers run on CUDA only; there is no CPU path")
        if batch.numel() % 4:
            raise ValueError("the state must hold a multiple of 4 elements")
        B = batch.shape[0]
        with torch.no_grad(), torch.cuda.device(dev):
            state = batch.to(device=dev, dtype=self.state_dtype, non_blocking=True).contiguous()
            if state.data_ptr() == batch.data_ptr():
                state = state.clone()
            tabs = VPStepTables(self.vp, ts.detach().to("cpu", torch.float64), n, bool(denoise),
                                float(eps), self._embedding())
            noise = None
            if self.noise is not None:
                noise = self.noise.to(device=dev, dtype=torch.float32).contiguous()
                if noise.shape[0] < n or tuple(noise.shape[1:]) != tuple(batch.shape):
                    raise ValueError(f"noise bank must be [{n},{tuple(batch.shape)}], got {tuple(noise.shape)}")
            net_in = state.to(torch.float32)
            tau32 = tabs.tau32.to(dev)
            sdt = L.dtype_code(self.state_dtype)
            stream = L.stream_ptr(dev)
            sp, ip, cnt = L.ptr(state), L.ptr(net_in), state.numel()
            record = None
            if self.record is not None:
                record = torch.empty(n, *state.shape, dtype=self.state_dtype, device=dev)
            for i in range(n + (1 if denoise else 0)):
                e = self.score_fn(net_in, tau32[i].expand(B)).to(torch.float32).contiguous()
                z = L.ptr(noise[i]) if (noise is not None and i < n) else None
                philox = 1 if (noise is None and i < n) else 0
                L.check(lib.psld_vp_em_update(sp, sp, sdt, ip, L.ptr(e), z, philox,
                                              C.byref(tabs.steps[i]), self.seed, i, cnt, stream),
                        "psld_vp_em_update")
                if i < n:
                    self._post_step(state, net_in, record, i, ts)
            if record is not None:
                self.record = record
            self._keep = (tabs, noise, net_in)
        return state

    def sample(self, batch, ts, n_discrete_steps, denoise=True, eps=1e-3):
        lib = L.lib()
        n = int(n_discrete_steps)
        self.nfe = n
        self._next_seed()
        if self.vp is not None:
            return self._sample_vp(batch, ts, n, denoise, eps)
        native = is_native(self.score_fn)
        if native:
            dev = next(self.score_fn.parameters()).device
        else:
            dev = batch.device if batch.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("psld_b200 samplers run on CUDA only; there is no CPU path")
        if batch.dim() != 4 or batch.shape[1] % 2:
            raise ValueError(f"expected a [B,2C,H,W] phase-space batch, got {tuple(batch.shape)}")
        B, C2, H, W = batch.shape
        chw = (C2 // 2) * H * W
        if chw % 4:
            raise ValueError("C*H*W must be a multiple of 4")
        with torch.no_grad(), torch.cuda.device(dev):
            state = batch.to(device=dev, dtype=self.state_dtype, non_blocking=True).contiguous()
            if state.data_ptr() == batch.data_ptr():
                state = state.clone()
            merged = (self.KIND == "sscs_sde" and self.noise is None and self.fuse_halves
                      and self.merge_noise and self.record is None and self.corrector_fn is None)
            tabs = StepTables(self.schedule, ts.detach().to("cpu", torch.float64), n, self.KIND,
                              bool(denoise), float(eps), self._embedding(), merge_noise=merged)
            self._merged = merged
            noise = None
            if self.noise is not None:
                noise = self.noise.to(device=dev, dtype=torch.float32).contiguous()
                need = (2 * n if self.KIND == "sscs_sde" else n)
                if noise.shape[0] < need or tuple(noise.shape[1:]) != (B, C2, H, W):
                    raise ValueError(f"noise bank must be [{need},{B},{C2},{H},{W}], got {tuple(noise.shape)}")
            record = None
            if self.record is not None:
                record = torch.empty(n, B, C2, H, W, dtype=self.state_dtype, device=dev)
            sdt = L.dtype_code(self.state_dtype)
            stream = L.stream_ptr(dev)
            if native and self.corrector_fn is None:
                self._run_native(lib, state, tabs, n, denoise, B, chw, noise, record, sdt, stream, dev)
            else:
                self._run_generic(lib, state, tabs, n, denoise, B, chw, noise, record, sdt, stream, dev,
                                  ts)
            if record is not None:
                self.record = record
        return state

    # ------------------------------------------------------------------ native loop
    def _run_native(self, lib, state, tabs, n, denoise, B, chw, noise, record, sdt, stream, dev):
        plan = self.score_fn.plan(B, 1, True)
        plan.x_in.copy_(state)                       # fp32 network input = f32(prior)
        table = tabs.time_table.to(dev)
        d = L.SamplerDesc()
        d.sampler = 0 if self.KIND == "sscs_sde" else 1
        d.n_steps, d.denoise, d.state_dtype = n, int(bool(denoise)), sdt
        d.fuse_halves = (2 if self._merged else 1) if (self.fuse_halves and record is None) else 0
        d.temb_op = plan.temb_op
        d.B, d.chw, d.seed = B, chw, self.seed
        d.state, d.net_in, d.eps = state.data_ptr(), plan.x_in.data_ptr(), plan.eps.data_ptr()
        d.time_table = table.data_ptr()
        d.noise = noise.data_ptr() if noise is not None else None
        if tabs.sscs is not None:
            d.sscs = C.cast(tabs.sscs, C.POINTER(L.SscsCoeffs))
        if tabs.em is not None:
            d.em = C.cast(tabs.em, C.POINTER(L.ScoreStep))
        if tabs.den is not None:
            d.den = C.pointer(tabs.den)
        d.record = record.data_ptr() if record is not None else None
        extra = None
        if self.use_graph and noise is None and record is None and n > 1:
            # CUDA-graph replay of one predictor step: per-step scalars live in device tables
            # indexed by a device-side step counter; needs a capturable (non-default) stream
            tab = tabs.sscs if tabs.sscs is not None else tabs.em
            raw = torch.frombuffer(bytearray(C.string_at(C.addressof(tab), C.sizeof(tab))),
                                   dtype=torch.uint8).to(dev)
            counter = torch.zeros(1, dtype=torch.int32, device=dev)
            if tabs.sscs is not None:
                d.sscs_dev = raw.data_ptr()
            else:
                d.em_dev = raw.data_ptr()
            d.step_counter = counter.data_ptr()
            extra = (raw, counter)
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                L.check(lib.psld_sampler_run(plan.op_array, plan.n_ops, C.byref(d),
                                             C.c_void_p(side.cuda_stream)), "psld_sampler_run")
            cur.wait_stream(side)
        else:
            L.check(lib.psld_sampler_run(plan.op_array, plan.n_ops, C.byref(d), stream),
                    "psld_sampler_run")
        self._keep = (table, noise, tabs, plan, extra)   # keep alive until the stream drains

    # ------------------------------------------------------------------ generic score_fn
    def _run_generic(self, lib, state, tabs, n, denoise, B, chw, noise, record, sdt, stream, dev,
                     ts):
        net_in = state.to(torch.float32)
        tau32 = tabs.tau32.to(dev)
        z = (lambda k: L.ptr(noise[k])) if noise is not None else (lambda k: None)
        sp, ip = L.ptr(state), L.ptr(net_in)

        def score(i):
            e = self.score_fn(net_in, tau32[i].expand(B))
            return e.to(torch.float32).contiguous()

        fuse = self.fuse_halves and record is None and self.corrector_fn is None
        if self.KIND == "sscs_sde":
            if fuse and n > 0:
                L.check(lib.psld_sscs_update(sp, sp, sdt, ip, None, z(0), None, None,
                                             C.byref(tabs.sscs[0]), L.STAGE_HALF_A, self.seed, 0, B,
                                             chw, stream), "psld_sscs_update")
            for i in range(n):
                if not fuse:
                    L.check(lib.psld_sscs_update(sp, sp, sdt, ip, None, z(2 * i), None, None,
                                                 C.byref(tabs.sscs[i]), L.STAGE_HALF_A, self.seed,
                                                 i, B, chw, stream), "psld_sscs_update")
                e = score(i)
                stages = L.STAGE_SCORE | L.STAGE_HALF_B
                if fuse and not self._merged and i + 1 < n:
                    stages |= L.STAGE_HALF_C
                zc = z(2 * i + 2) if (stages & L.STAGE_HALF_C) else None
                L.check(lib.psld_sscs_update(sp, sp, sdt, ip, L.ptr(e), None, z(2 * i + 1), zc,
                                             C.byref(tabs.sscs[i]), stages, self.seed, i, B, chw,
                                             stream), "psld_sscs_update")
                self._post_step(state, net_in, record, i, ts)
        else:
            for i in range(n):
                e = score(i)
                L.check(lib.psld_em_update(sp, sp, sdt, ip, L.ptr(e), z(i), 0 if noise is not None else 1,
                                           C.byref(tabs.em[i]), self.seed, i, B, chw, stream),
                        "psld_em_update")
                self._post_step(state, net_in, record, i, ts)
        if denoise:
            e = score(n)
            L.check(lib.psld_em_update(sp, sp, sdt, ip, L.ptr(e), None, 0, C.byref(tabs.den),
                                       self.seed, n, B, chw, stream), "psld_em_update")

    def _post_step(self, state, net_in, record, i, ts):
        if self.corrector_fn is not None:
            new, _ = self.corrector_update_fn(state, ts[i], ts[i + 1] - ts[i])
            state.copy_(new)
            net_in.copy_(state)
        if record is not None:
            record[i].copy_(state)


@register_module(category="samplers", name="sscs_sde_b200")
class SSCSSampler(_FusedSampler):
    """Symmetric-splitting sampler for PSLD (reference sde.py:227-370)."""
    KIND = "sscs_sde"


@register_module(category="samplers", name="em_sde_b200")
class EulerMaruyamaSampler(_FusedSampler):
    """Euler-Maruyama sampler (reference sde.py:8-58)."""
    KIND = "em_sde"


@register_module(category="samplers", name="cc_em_sde_b200")
class ClassCondEulerMaruyamaSampler(_FusedSampler):
    """Classifier-guided Euler-Maruyama sampler (reference ``ClassCondEulerMaruyamaSampler``,
    sde.py:61-122): ``cls(config, sde, score_fn, clf_fn, corrector_fn=None)``; every predictor step adds
    ``g^2 * clf_temp * d/du log p(y | u, t)`` to the reverse drift, with
    ``y = config.clf.evaluation.label_to_sample`` and ``clf_temp = config.clf.evaluation.clf_temp``.

    The classifier is the caller's differentiable module (the reference's ``NCSNppClassifier`` is
    outside the hot-path scope); its input gradient is taken with
    ``torch.autograd.grad`` exactly as the reference does (sde.py:82-90: float32 input, REVERSE time
    ``t`` as float32, ``log_softmax`` then the selected class).  The score network call and the guided
    update (``psld_em_update_guided``: drift + guidance + diffusion in one pass) run on this library.
    The denoising call keeps the guided MEAN of one more predictor step (sde.py:112-117)."""
    KIND = "em_sde"

    def __init__(self, config, sde, score_fn, clf_fn, corrector_fn=None):
        super().__init__(config, sde, score_fn, corrector_fn=corrector_fn)
        if self.vp is not None:
            raise ValueError("cc_em_sde is defined for the PSLD SDE")
        self.clf_fn = clf_fn
        ev = config.clf.evaluation
        self.y = ev.label_to_sample
        self.clf_temp = float(ev.clf_temp)

    def _guidance(self, state, t):
        """d/du log p(y | u, t) as fp32 (sde.py:82-90)."""
        B = state.shape[0]
        with torch.inference_mode(False), torch.enable_grad():
            x_in = state.detach().clone().requires_grad_()
            tt = torch.full((B,), float(t), dtype=torch.float32, device=state.device)
            logits = self.clf_fn(x_in.to(torch.float32), tt)
            log_probs = torch.nn.functional.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), self.y]
            grad = torch.autograd.grad(selected.sum(), x_in)[0]
        return grad.detach().to(torch.float32).contiguous()

    def sample(self, batch, ts, n_discrete_steps, denoise=True, eps=1e-3):
        lib = L.lib()
        n = int(n_discrete_steps)
        self.nfe = n
        self._next_seed()
        if is_native(self.score_fn):
            dev = next(self.score_fn.parameters()).device
        else:
            dev = batch.device if batch.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("psld_b200 samplers run on CUDA only; there is no CPU path")
        if batch.dim() != 4 or batch.shape[1] % 2:
            raise ValueError(f"expected a [B,2C,H,W] phase-space batch, got {tuple(batch.shape)}")
        B, C2, H, W = batch.shape
        chw = (C2 // 2) * H * W
        if chw % 4:
            raise ValueError("C*H*W must be a multiple of 4")
        with torch.no_grad(), torch.cuda.device(dev):
            state = batch.to(device=dev, dtype=self.state_dtype).contiguous().clone()
            ts64 = ts.detach().to("cpu", torch.float64)
            tabs = StepTables(self.schedule, ts64, n, "em_sde", bool(denoise), float(eps), self._embedding())
            noise = None
            if self.noise is not None:
                noise = self.noise.to(device=dev, dtype=torch.float32).contiguous()
                if noise.shape[0] < n or tuple(noise.shape[1:]) != (B, C2, H, W):
                    raise ValueError(f"noise bank must be [{n},{B},{C2},{H},{W}], got {tuple(noise.shape)}")
            record = None
            if self.record is not None:
                record = torch.empty(n, B, C2, H, W, dtype=self.state_dtype, device=dev)
            net_in = state.to(torch.float32)
            tau32 = tabs.tau32.to(dev)
            sdt = L.dtype_code(self.state_dtype)
            stream = L.stream_ptr(dev)
            sp, ip = L.ptr(state), L.ptr(net_in)
            # reverse times the classifier sees: ts[i], then fl32(T - eps) for the denoising call
            t_den = float(torch.tensor(self.schedule.T - float(eps), dtype=torch.float32))
            for i in range(n + (1 if denoise else 0)):
                e = self.score_fn(net_in, tau32[i].expand(B)).to(torch.float32).contiguous()
                g = self._guidance(state, float(ts64[i]) if i < n else t_den)
                last = i == n
                z = L.ptr(noise[i]) if (noise is not None and not last) else None
                philox = 1 if (noise is None and not last) else 0
                co = tabs.em[i] if not last else tabs.den
                L.check(lib.psld_em_update_guided(sp, sp, sdt, ip, L.ptr(e), L.ptr(g), self.clf_temp, z,
                                                  philox, C.byref(co), self.seed, i, B, chw, stream),
                        "psld_em_update_guided")
                if not last:
                    self._post_step(state, net_in, record, i, ts)
            if record is not None:
                self.record = record
            self._keep = (tabs, noise, net_in)
        return state


@register_module(category="samplers", name="ip_em_sde_b200")
class InpaintEulerMaruyamaSampler(_FusedSampler):
    """Euler-Maruyama inpainting sampler (reference ``ES3EulerMaruyamaInpainter``,
    sde.py:125-224): ``sample((x_0, mask), ts, n, denoise, eps)``; after every predictor step
    the known region (``mask == 1``) is replaced by a fresh perturbation of ``x_0`` at the
    current noise level (one fused ``psld_inpaint_combine`` pass).

    ``self.prior`` (optional ``[B,2C,H,W]``) replaces the device-drawn prior; ``self.noise``
    (optional, parity mode) is a dict of pre-drawn N(0,1) tensors in the reference's draw
    order: ``pred [n(+1),B,2C,H,W]``, ``m0 [n+2,B,C,H,W]``, ``eps [n+2,B,2C,H,W]`` (index 0 =
    initial latent, 1..n = steps, n+1 = denoise call; the denoise entries may be omitted when
    ``denoise`` is False).  HSM / DSM follows ``config.training.mode`` (sde.py:137-143)."""
    KIND = "em_sde"

    def __init__(self, config, sde, score_fn, corrector_fn=None):
        super().__init__(config, sde, score_fn, corrector_fn=corrector_fn)
        self.prior = None
        tr = getattr(config, "training", None)
        mode = tr.get("mode", "hsm") if hasattr(tr, "get") else getattr(tr, "mode", "hsm")
        self.hsm = str(mode) == "hsm"

    def sample(self, batch, ts, n_discrete_steps, denoise=True, eps=1e-3):
        lib = L.lib()
        x_0, mask = batch
        n = int(n_discrete_steps)
        self.nfe = n
        self._next_seed()
        if self.record is not None or self.corrector_fn is not None:
            raise NotImplementedError("ip_em_sde_b200 does not support `record` / a custom corrector_fn "
                                      "(the reference's inpainting sampler has neither)")
        if is_native(self.score_fn):
            dev = next(self.score_fn.parameters()).device
        else:
            dev = x_0.device if x_0.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("psld_b200 samplers run on CUDA only; there is no CPU path")
        if x_0.dim() != 4:
            raise ValueError(f"expected x_0 [B,C,H,W], got {tuple(x_0.shape)}")
        B, Cc, H, W = x_0.shape
        chw = Cc * H * W
        if chw % 4:
            raise ValueError("C*H*W must be a multiple of 4")
        with torch.no_grad(), torch.cuda.device(dev):
            x0 = x_0.to(device=dev, dtype=torch.float32).contiguous()
            mk = mask.to(device=dev, dtype=torch.float32).expand(B, Cc, H, W).contiguous()
            stream = L.stream_ptr(dev)
            if self.prior is not None:
                state = self.prior.to(device=dev, dtype=self.state_dtype).contiguous().clone()
            else:
                pr = torch.empty(B, 2 * Cc, H, W, dtype=torch.float32, device=dev)
                L.check(lib.psld_prior_sample(L.ptr(pr), float(self.schedule.m) ** 0.5, self.seed, B,
                                              chw, stream), "psld_prior_sample")
                state = pr.to(self.state_dtype)
            tabs = StepTables(self.schedule, ts.detach().to("cpu", torch.float64), n, "em_sde",
                              bool(denoise), float(eps), self._embedding())
            ip = InpaintTables(self.schedule, ts.detach().to("cpu", torch.float64), n, bool(denoise),
                               float(eps), self.hsm)
            bank = None
            if self.noise is not None:
                bank = {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in self.noise.items()}
            net_in = torch.empty(B, 2 * Cc, H, W, dtype=torch.float32, device=dev)
            sdt = L.dtype_code(self.state_dtype)
            sp, np_ = L.ptr(state), L.ptr(net_in)
            tau32 = tabs.tau32.to(dev)

            def combine(k):
                zm = L.ptr(bank["m0"][k]) if bank is not None and not self.hsm else None
                ze = L.ptr(bank["eps"][k]) if bank is not None else None
                L.check(lib.psld_inpaint_combine(sp, sdt, np_, L.ptr(x0), L.ptr(mk), zm, ze,
                                                 C.byref(ip.steps[k]), self.seed, k, B, chw, stream),
                        "psld_inpaint_combine")

            def score(i):
                e = self.score_fn(net_in, tau32[i].expand(B))
                return e.to(torch.float32).contiguous()

            combine(0)
            for i in range(n):
                e = score(i)
                z = L.ptr(bank["pred"][i]) if bank is not None else None
                L.check(lib.psld_em_update(sp, sp, sdt, np_, L.ptr(e), z, 0 if bank is not None else 1,
                                           C.byref(tabs.em[i]), self.seed, i, B, chw, stream),
                        "psld_em_update")
                combine(i + 1)
            if denoise:
                e = score(n)
                L.check(lib.psld_em_update(sp, sp, sdt, np_, L.ptr(e), None, 0, C.byref(tabs.den),
                                           self.seed, n, B, chw, stream), "psld_em_update")
                combine(n + 1)
            self._keep = (x0, mk, bank, net_in, tabs, ip)
        return state
