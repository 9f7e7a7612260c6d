"""Black-box probability-flow ODE sampler on the fused sm_100a kernels.

Drop-in for the reference's ``BBODESampler`` (``main/samplers/ode.py:8-76``; registry name
``bb_ode``, used by ``scripts_psld/**/sample_uncond_psld_ode.sh`` and, over the VP-SDE baseline, by
``scripts_psld/ablations/uncond/cifar10/sample_uncond_vpsde_ode.sh`` with
``evaluation.sampler.{solver=RK45, rtol, atol}``):

  ``cls(config, sde, score_fn, corrector_fn=None)``
  ``.sample(batch, ts, n_discrete_steps, denoise=True, eps=1e-3)``; ``.nfe`` counts score_fn calls,
  ``.n_steps`` / ``.mean_nfe`` as in the reference.

What the reference does: ``torchdiffeq.odeint(ode_fn, x, [0, T - eps], rtol, atol,
method="scipy_solver", options={"solver": "RK45"})`` with ``ode_fn(t, x) = reverse_sde(x, t, score_fn,
probability_flow=True)[0]`` (``psld.py:345-364``).  torchdiffeq (0.2.3 in the reference's Pipfile) is a
third-party dependency that is NOT part of the reference tree; its ``scipy_solver`` is a thin wrapper
around ``scipy.integrate.solve_ivp``: the state lives in a flat float64 numpy vector on the HOST, and
every function evaluation copies it to the device (rounded to the batch dtype, as is ``t``), calls
``ode_fn`` and copies the float64 drift back.

Here the same algorithm - scipy's ``RK45`` (Dormand-Prince 5(4), ``select_initial_step``, the
``RungeKutta._step_impl`` accept / reject and step-size rules with SAFETY 0.9, factors in [0.2, 10],
error exponent -1/5, error norm = RMS of err / (atol + rtol max(|y|, |y_new|))) - is restated with the
state, the seven stage derivatives and every vector operation resident on the GPU
(``psld_rk_combine`` / ``psld_rk_error`` / ``psld_reverse_drift``); the host only sees one scalar
(the error norm) per attempted step.  ``solver`` must be ``RK45`` (what every shipped script uses).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from .guidance import is_native
from .registry import register_module
from .samplers import Sampler
from .schedule import PSLDSchedule, VPSchedule, _fill_score, _score_rows

# Dormand-Prince 5(4) tableau (scipy.integrate._ivp.rk.RK45)
_C = [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0]
_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
]
_B = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]
_E = [-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40]
_SAFETY, _MIN_FACTOR, _MAX_FACTOR = 0.9, 0.2, 10.0
_ORDER = 4                          # error_estimator_order of RK45


def _darr(v):
    return (C.c_double * len(v))(*v)


@register_module(category="samplers", name="bb_ode_b200")
class BBODESampler(Sampler):
    """Probability-flow ODE sampler (reference ode.py:8-76)."""

    def __init__(self, config, sde, score_fn, corrector_fn=None):
        super().__init__(config, sde, score_fn, corrector_fn=corrector_fn)
        # the VP-SDE baseline (state [B,C,H,W], vpsde.py:48-67) or PSLD (phase-space state [B,2C,H,W])
        self.vp = isinstance(sde, VPSchedule) or str(getattr(sde, "type", "")) == "vpsde"
        if self.vp:
            self.schedule = sde if isinstance(sde, VPSchedule) else VPSchedule.from_sde(sde)
        else:
            self.schedule = sde if isinstance(sde, PSLDSchedule) else PSLDSchedule.from_sde(sde)
        self.nfe = 0
        s = config.evaluation.sampler
        get = (lambda k: s.get(k)) if hasattr(s, "get") else (lambda k: getattr(s, k))
        self.rtol = float(get("rtol"))
        self.atol = float(get("atol"))
        self.solver_opts = {"solver": get("solver")}
        if str(self.solver_opts["solver"]).upper() != "RK45":
            raise NotImplementedError("bb_ode_b200 restates scipy's RK45 driver only (the solver every "
                                      "shipped sampling script selects)")
        self._counter = 0
        self.steps_accepted = self.steps_rejected = 0

    @property
    def n_steps(self):
        return self.nfe

    @property
    def mean_nfe(self):
        if self._counter != 0:
            return self.nfe / self._counter
        raise ValueError("Run .sample() to compute mean_nfe")

    def predictor_update_fn(self, x, t, dt):
        pass

    # ------------------------------------------------------------------ one ode_fn evaluation
    def _embedding(self):
        return getattr(self.score_fn, "embedding_type", "fourier")

    def _drift(self, ctx, t, u_state, net_in, out):
        """out (f64) = reverse_sde(u, t, score_fn, probability_flow=True)[0]; ``u_state`` is the state
        in the batch dtype, ``net_in`` its float32 view (the same buffer for a float32 batch)."""
        lib, B, chw, stream, sdt = ctx["lib"], ctx["B"], ctx["chw"], ctx["stream"], ctx["sdt"]
        self.nfe += 1
        # torchdiffeq hands t over in the batch dtype (convert_func_to_numpy: torch.tensor(t).to(dtype))
        t_eff = float(np.float32(t)) if ctx["batch_f32"] else float(t)
        tau = torch.tensor([self.schedule.T - t_eff], dtype=torch.float64)
        if self.vp:
            beta = self.schedule.beta_t(tau)
            co = L.VpStep()
            co.half_beta = float(0.5 * beta[0])
            co.g2 = float(torch.sqrt(beta)[0] ** 2)
            co.neg_inv_std = float(-1.0 / self.schedule.std(tau)[0])
        else:
            rows = _score_rows(self.schedule, tau, torch.ones(1, dtype=torch.float64))
            co = L.ScoreStep()
            _fill_score(co, self.schedule, rows, 0)
        tau32 = tau.to(torch.float32)
        plan = ctx["plan"]
        if plan is not None:          # native network: net_in IS plan.x_in
            plan.time_buf.copy_(torch.log(tau32) if self._embedding() == "fourier" else tau32)
            plan.run(stream)
            e = plan.eps
        else:
            e = self.score_fn(net_in, tau32.to(net_in.device).expand(B)).to(torch.float32).contiguous()
        if self.vp:
            L.check(lib.psld_vp_reverse_drift(L.ptr(out), L.ptr(u_state), sdt, L.ptr(e), C.byref(co), 0.5,
                                              u_state.numel(), stream), "psld_vp_reverse_drift")
        else:
            L.check(lib.psld_reverse_drift(L.ptr(out), L.ptr(u_state), sdt, L.ptr(e), C.byref(co), 0.5, B, chw,
                                           stream), "psld_reverse_drift")
        ctx["keep"] = (e, co)

    def sample(self, batch, ts, n_discrete_steps, denoise=True, eps=1e-3):
        lib = L.lib()
        native = is_native(self.score_fn)
        if native:
            dev = next(self.score_fn.parameters()).device
        else:
            dev = batch.device if batch.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("psld_b200 samplers run on CUDA only; there is no CPU path")
        if batch.dim() != 4 or (batch.shape[1] % 2 and not self.vp):
            raise ValueError(f"expected a [B,2C,H,W] phase-space batch, got {tuple(batch.shape)}")
        B, C2, H, W = batch.shape                       # C2 = the state's channel count (VP-SDE: C)
        chw = (C2 // 2) * H * W
        if (batch.numel() if self.vp else chw) % 4:
            raise ValueError("C*H*W must be a multiple of 4")
        self._counter += 1
        n = batch.numel()
        batch_f32 = batch.dtype != torch.float64
        rtol, atol = max(self.rtol, 100 * np.finfo(np.float64).eps), self.atol      # scipy validate_tol
        with torch.no_grad(), torch.cuda.device(dev):
            stream = L.stream_ptr(dev)
            plan = self.score_fn.plan(B, 1, True) if native else None
            f64 = dict(dtype=torch.float64, device=dev)
            y = batch.to(**f64).contiguous().clone()            # scipy's float64 state vector
            y_new = torch.empty_like(y)
            K = torch.empty(7, n, **f64)
            err_sum = torch.zeros(1, **f64)
            # the state as ode_fn sees it: rounded to the batch dtype; its float32 view feeds the network
            net_in = plan.x_in if native else torch.empty(B, C2, H, W, dtype=torch.float32, device=dev)
            u64 = None if batch_f32 else torch.empty_like(y)
            ctx = dict(lib=lib, B=B, chw=chw, stream=stream, plan=plan, batch_f32=batch_f32,
                       sdt=L.F32 if batch_f32 else L.F64)

            def stage(src, coef, h, out64=None):
                """(u_state, net_in) <- round(src + h * sum_j coef[j] K[j]); optionally also float64."""
                o64 = out64 if out64 is not None else u64
                L.check(lib.psld_rk_combine(L.ptr(o64), L.ptr(net_in), L.ptr(src), L.ptr(K), _darr(coef or [0.0]),
                                            len(coef), h, n, stream), "psld_rk_combine")
                return net_in if batch_f32 else o64

            def fun(t, u_state, out):
                self._drift(ctx, t, u_state, net_in, out)

            def rms(v):                        # scipy.integrate._ivp.common.norm
                return float(torch.linalg.vector_norm(v)) / math.sqrt(n)

            t0, t_bound = 0.0, float(self.schedule.T - eps)
            # ---- RungeKutta.__init__: f0 and select_initial_step (common.py)
            fun(t0, stage(y, [], 0.0), K[0])
            f0 = K[0]
            scale = atol + y.abs() * rtol
            d0, d1 = rms(y.reshape(-1) / scale.reshape(-1)), rms(f0 / scale.reshape(-1))
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            h0 = min(h0, abs(t_bound - t0))
            fun(t0 + h0, stage(y, [1.0], h0), K[1])             # y1 = y0 + h0 * f0 (K[0] only)
            d2 = rms((K[1] - f0) / scale.reshape(-1)) / h0
            h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / (_ORDER + 1))
            h_abs = min(100 * h0, h1, abs(t_bound - t0))
            del scale
            # ---- solve_ivp loop: solver.step() until t reaches t_bound
            t = t0
            exponent = -1.0 / (_ORDER + 1)
            self.steps_accepted = self.steps_rejected = 0
            while t < t_bound:
                min_step = 10 * abs(np.nextafter(t, np.inf) - t)
                h_abs = max(h_abs, min_step)
                rejected = False
                while True:
                    if h_abs < min_step:
                        raise RuntimeError("bb_ode: required step size is less than spacing between numbers")
                    t_new = min(t + h_abs, t_bound)
                    h = t_new - t
                    h_abs = abs(h)
                    for s in range(1, 6):                        # rk_step: K[1..5]
                        fun(t + _C[s] * h, stage(y, _A[s], h), K[s])
                    u_new = stage(y, _B, h, out64=y_new)         # y_new = y + h * K[:6].B
                    fun(t + h, u_new, K[6])                      # f_new
                    err_sum.zero_()
                    L.check(lib.psld_rk_error(L.ptr(y), L.ptr(y_new), L.ptr(K), _darr(_E), 7, h, atol, rtol, n,
                                              L.ptr(err_sum), stream), "psld_rk_error")
                    error_norm = math.sqrt(float(err_sum) / n)
                    if error_norm < 1:
                        factor = _MAX_FACTOR if error_norm == 0 else min(_MAX_FACTOR, _SAFETY * error_norm ** exponent)
                        if rejected:
                            factor = min(1.0, factor)
                        h_abs *= factor
                        self.steps_accepted += 1
                        break
                    h_abs *= max(_MIN_FACTOR, _SAFETY * error_norm ** exponent)
                    rejected = True
                    self.steps_rejected += 1
                t = t_new
                y, y_new = y_new, y
                K[0].copy_(K[6])                                 # FSAL: f of the accepted point
            # odeint returns the solution in the batch dtype (ScipyWrapperODESolver.integrate)
            x = y.to(batch.dtype) if batch_f32 else y
            if denoise:
                # denoise_fn(x, T - eps, eps) = x + f * eps with t a float64 vector (ode.py:36-39,66-75)
                ctx["batch_f32"] = False                         # t is NOT rounded here
                net_in.copy_(x)
                u_state = x.to(torch.float32).contiguous() if batch_f32 else x.contiguous()
                fbar = K[1].view(B, C2, H, W)
                self._drift(ctx, self.schedule.T - eps, u_state, net_in, fbar)
                x = x.to(torch.float64) + fbar * float(eps)
        return x
