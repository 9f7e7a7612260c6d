"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of PSLD's reverse-time sampling path.

This file is the checker for the CUDA path in ``psld_b200``; it is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  The product never routes through it.

It restates, in plain float64 scalar algebra + torch-CPU tensor ops, the reference
functions of SURVEY.md §8(a).  All ``file:line`` citations are relative to the
reference repo (mandt-lab/PSLD, ``/root/reference``):

  * schedule / covariance / Cholesky scalars ... ``main/models/sde/psld.py:38-44,86-220``
  * SSCS half-step mean / variance ............ ``main/samplers/sde.py:236-312``
  * SSCS score step ........................... ``main/samplers/sde.py:314-329``
  * score from eps (fp32 coefficient cast) .... ``main/models/sde/psld.py:230-260``
  * forward / reverse drift, EM step, denoise . ``main/models/sde/psld.py:330-364``,
                                                ``main/samplers/sde.py:16-58,338-370``
  * classifier-guided EM (cc_em_sde) .......... ``main/samplers/sde.py:61-122``
  * probability-flow ODE sampler (bb_ode) ..... ``main/samplers/ode.py:8-76`` (solver: scipy RK45, as
                                                called by torchdiffeq 0.2.3's scipy_solver wrapper)
  * time grid ................................. ``main/models/wrapper.py:51-54,101-114``
  * NCSN++ forward ............................ ``main/models/score_fn/song_sde/ncsnpp.py:287-438``
  * ResnetBlockBigGANpp / AttnBlockpp / NIN ... ``.../layerspp.py:75-91,242-274``, ``.../layers.py:531-540``
  * FIR resampling (upfirdn2d) ................ ``.../up_or_down_sampling.py:144-257``,
                                                ``.../op/upfirdn2d.py:159-200``

Third-party arithmetic (conv2d / group_norm / softmax / linear) is PyTorch ATen on
CPU in fp32, exactly what the reference itself calls (``torch==1.13.1`` pinned in the
reference ``Pipfile:7``; 2.11.0 here).

PARITY PINNING: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference itself run in the build container
(``oracle/make_golden.py`` → ``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py``)
and, when ``/root/reference`` is present, directly (``tests/test_oracle_vs_reference.py``).

Known, deliberate deviation: the reference's very first SSCS half-step multiplies an
fp32 state by a python scalar before promoting to fp64 (``sde.py:249`` with the fp32
prior); this oracle promotes the prior to fp64 first.  Effect ≤ 1e-10 relative.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# 1. SDE scalars (float64)
# --------------------------------------------------------------------------------------
class PSLDScalars:
    """Scalar (per-time) restatement of ``PSLD`` (``psld.py:14-60``)."""

    def __init__(self, config):
        c = config.model.sde
        self.beta_0, self.beta_1 = float(c.beta_min), float(c.beta_max)
        self.nu, self.gamma = float(c.nu), float(c.gamma)
        assert self.nu != 0 or self.gamma != 0                     # psld.py:24
        self.m_inv = (self.gamma - self.nu) ** 2 / 4               # psld.py:25
        self.m = 1 / self.m_inv
        self.kappa = float(c.kappa)
        self.mm_0 = self.kappa * self.m                            # psld.py:30
        self.eps = float(c.numerical_eps)
        self.decomp_mode = c.decomp_mode
        assert self.decomp_mode in ("lower", "upper")              # psld.py:33
        self.T = 1.0

    @property
    def mode(self):                                                # psld.py:50-56
        if self.gamma == 0:
            return "score_m"
        if self.nu == 0:
            return "score_x"
        return "score_xm"

    def beta_t(self, t):                                           # psld.py:38-40
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def b_t(self, t):                                              # psld.py:42-44
        return self.beta_0 * t + 0.5 * (t ** 2) * (self.beta_1 - self.beta_0)

    def cov(self, xx_0, mm_0, t):                                  # psld.py:86-152
        nu, ga, mi, m = self.nu, self.gamma, self.m_inv, self.m
        lam = (nu + ga) / 2
        b = self.b_t(t)
        b2 = b ** 2
        s, si = math.exp(-lam * b), math.exp(lam * b)
        xx = (mi / 4 * b2 * xx_0 + mi ** 2 / 4 * b2 * mm_0 + (nu - ga) / 2 * b * xx_0
              + (-mi / 2) * b2 + (ga - nu) / 2 * b + (si - 1) + xx_0) * s
        xm = ((ga - nu) / 8 * b2 * xx_0 + mi * (ga - nu) / 8 * b2 * mm_0 + (-1 / 2) * b * xx_0
              + mi / 2 * b * mm_0 + (nu - ga) / 4 * b2) * s
        mm = (1 / 4 * b2 * xx_0 + mi / 4 * b2 * mm_0 + (ga - nu) / 2 * b * mm_0
              + (-1 / 2) * b2 + m * (nu - ga) / 2 * b + m * (si - 1) + mm_0) * s
        return xx + self.eps, xm, mm + self.eps

    def get_coeff(self, var):                                      # psld.py:154-186
        xx, xm, mm = var
        if self.decomp_mode == "lower":
            l11 = _sqrt(xx); l21 = xm / l11; l22 = _sqrt(mm - l21 ** 2.0)
            out = (l11, 0.0, l21, l22)
        else:
            u22 = _sqrt(mm); u12 = xm / u22; u11 = _sqrt(xx - u12 ** 2.0)
            out = (u11, u12, 0.0, u22)
        if any(math.isnan(v) for v in out):
            raise ValueError("Numerical precision error.")         # psld.py:171
        return out

    def get_inv_coeff(self, var):                                  # psld.py:188-220
        xx, xm, mm = var
        det = xx * mm - xm ** 2
        if self.decomp_mode == "lower":
            out = (_sqrt(1 / xx), -xm / (_sqrt(xx) * _sqrt(det)), 0.0, _sqrt(xx / det))
        else:
            out = (_sqrt(mm / det), 0.0, -xm / (_sqrt(mm) * _sqrt(det)), _sqrt(1 / mm))
        if any(math.isnan(v) for v in out):
            raise ValueError("Numerical precision error.")
        return out

    # ---- SSCS half step (sde.py:236-292) ----
    def half_step(self, t, h):
        """Returns (a_xx, a_xm, a_mx, a_mm), (c11, c12, c21, c22): u' = A u + C z."""
        nu, ga = self.nu, self.gamma
        db = self.b_t(self.T - (t + h)) - self.b_t(self.T - t)     # sde.py:241
        s = math.exp((nu + ga) / 4 * db)
        A_1 = (nu - ga) / 4
        A_2 = -((ga - nu) ** 2) / 8
        C_2 = (ga - nu) / 4
        a = (s * (1 - A_1 * db), s * (-A_2 * db), s * (-0.5 * db), s * (1 - C_2 * db))
        lam = (nu + ga) / 2
        E, Ei = math.exp(lam * db), math.exp(-lam * db)
        db2 = db ** 2
        xx = (-self.m_inv / 2 * db2 - (ga - nu) / 2 * db + (Ei - 1)) * E + self.eps
        xm = ((ga - nu) / 4 * db2) * E
        mm = (-0.5 * db2 - self.m * (nu - ga) / 2 * db + self.m * (Ei - 1)) * E + self.eps
        return a, self.get_coeff((xx, xm, mm))


def _sqrt(v):
    return math.sqrt(v) if v >= 0 else float("nan")


def time_grid(config, T=1.0):
    """``wrapper.py:51-54,101-114``: returns (ts float64 ndarray [n+1], n)."""
    ev = config.evaluation
    n = ev.n_discrete_steps - 1 if ev.denoise else ev.n_discrete_steps
    t_final = T - ev.eval_eps
    ts = np.linspace(0.0, t_final, n + 1, dtype=np.float64)
    if ev.stride_type == "quadratic":
        ts = t_final * (1 - (ts / t_final) ** 2.0)[::-1].copy()
    return ts, n


def prior_sampling(sde: PSLDScalars, shape, generator=None):
    """``psld.py:366-370``."""
    p_x = torch.randn(*shape, generator=generator)
    p_m = torch.randn(*shape, generator=generator) * np.sqrt(sde.m)
    return torch.cat([p_x, p_m], dim=1)


# --------------------------------------------------------------------------------------
# 2. Samplers (state float64, network I/O float32 — SURVEY.md §0.2)
# --------------------------------------------------------------------------------------
def score_from_eps(sde: PSLDScalars, eps32: torch.Tensor, tau: float):
    """``psld.py:230-260``: score = -L^{-T} eps with the coefficients rounded to fp32."""
    c11, c12, c21, c22 = sde.get_inv_coeff(sde.cov(0.0, sde.mm_0, tau))
    f32 = lambda v: torch.tensor(v, dtype=torch.float64).to(torch.float32)
    if sde.decomp_mode == "lower" and sde.mode == "score_m":
        return torch.cat([torch.zeros_like(eps32), -f32(c22) * eps32], dim=1)
    if sde.decomp_mode == "upper" and sde.mode == "score_x":
        return torch.cat([-f32(c11) * eps32, torch.zeros_like(eps32)], dim=1)
    ex, em = torch.chunk(eps32, 2, dim=1)
    sx = -f32(c11) * ex - f32(c12) * em
    sm = -f32(c21) * ex - f32(c22) * em
    return torch.cat([sx, sm], dim=1)


def _tvec(u, val):
    return torch.full((u.shape[0],), val, dtype=torch.float32)


def reverse_drift(sde: PSLDScalars, score_fn, u: torch.Tensor, t: float):
    """``psld.py:330-364`` (probability_flow=False): returns (f_bar, (g_x, g_m))."""
    tau = sde.T - t
    x, m = torch.chunk(u, 2, dim=1)
    beta = sde.beta_t(tau)
    fx = 0.5 * beta * (sde.m_inv * m - sde.gamma * x)
    fm = 0.5 * beta * (-sde.nu * m - x)
    gx, gm = math.sqrt(beta * sde.gamma), math.sqrt(beta * sde.m * sde.nu)
    tau32 = float(np.float32(tau))
    eps = score_fn(u.to(torch.float32), _tvec(u, tau32))
    score = score_from_eps(sde, eps, tau)
    sx, sm = torch.chunk(score, 2, dim=1)
    fbar = torch.cat([-fx + gx ** 2 * sx, -fm + gm ** 2 * sm], dim=1)
    return fbar, (gx, gm)


def _denoise(sde, score_fn, u, eps):
    """``denoising_fn`` as called from ``sample`` (sde.py:52-57,364-369): the reference builds
    ``t = torch.tensor(T - eps)`` and ``dt = torch.tensor(eps)`` as **float32** tensors, so the
    denoise step runs at t = fl32(T - eps) with step fl32(eps)."""
    t_den = float(np.float32(sde.T - eps))
    dt_den = float(np.float32(eps))
    fbar, _ = reverse_drift(sde, score_fn, u, t_den)
    return u + fbar * dt_den


def em_sample(config, score_fn, u0, ts, n, noise, denoise=True, eps=1e-3, record=None):
    """``EulerMaruyamaSampler.sample`` (``sde.py:38-58``).  ``noise``: n tensors [B,2C,H,W]."""
    sde = PSLDScalars(config)
    u = u0.to(torch.float64)
    C = u.shape[1] // 2
    with torch.no_grad():
        for i in range(n):
            t, dt = float(ts[i]), float(ts[i + 1] - ts[i])
            fbar, (gx, gm) = reverse_drift(sde, score_fn, u, t)
            z = noise[i].to(torch.float64)
            g = torch.cat([torch.full_like(u[:, :C], gx), torch.full_like(u[:, C:], gm)], dim=1)
            u = (u + fbar * dt) + g * math.sqrt(dt) * z            # sde.py:23-25
            if record is not None:
                record(i, u)
        if denoise:                                                # sde.py:28-36,52-57
            u = _denoise(sde, score_fn, u, eps)
    return u


def cc_em_sample(config, score_fn, clf_fn, u0, ts, n, noise, y, clf_temp, denoise=True, eps=1e-3,
                 record=None):
    """``ClassCondEulerMaruyamaSampler.sample`` (``sde.py:61-122``): Euler-Maruyama whose reverse
    drift gets the classifier-guidance term ``g^2 * clf_temp * d/du log p(y | u, t)``
    (``sde.py:82-93``; the classifier sees the REVERSE time ``t`` as float32, ``sde.py:84-87``).  The
    denoising call is a predictor step at fl32(T - eps) whose MEAN (guidance included) is kept
    (``sde.py:112-117``); its noise draw is discarded.  ``noise``: n tensors [B,2C,H,W]."""
    sde = PSLDScalars(config)
    u = u0.to(torch.float64)
    C = u.shape[1] // 2
    idx = torch.as_tensor(y).expand(u.shape[0]) if not torch.is_tensor(y) or y.dim() == 0 else y

    def guided_drift(u, t):
        fbar, (gx, gm) = reverse_drift(sde, score_fn, u, t)
        with torch.enable_grad():
            x_in = u.clone().requires_grad_()
            logits = clf_fn(x_in.to(torch.float32), _tvec(u, float(np.float32(t))))
            sel = F.log_softmax(logits, dim=-1)[range(len(logits)), idx]
            grad = torch.autograd.grad(sel.sum(), x_in)[0] * clf_temp
        g2 = torch.cat([torch.full_like(u[:, :C], gx ** 2), torch.full_like(u[:, C:], gm ** 2)], dim=1)
        return fbar + g2 * grad, (gx, gm)

    with torch.no_grad():
        for i in range(n):
            t, dt = float(ts[i]), float(ts[i + 1] - ts[i])
            f, (gx, gm) = guided_drift(u, t)
            g = torch.cat([torch.full_like(u[:, :C], gx), torch.full_like(u[:, C:], gm)], dim=1)
            u = (u + f * dt) + g * math.sqrt(dt) * noise[i].to(torch.float64)
            if record is not None:
                record(i, u)
        if denoise:
            f, _ = guided_drift(u, float(np.float32(sde.T - eps)))
            u = u + f * float(np.float32(eps))
    return u


def pflow_drift(sde: PSLDScalars, score_fn, u: torch.Tensor, t: float):
    """``reverse_sde(u, t, score_fn, probability_flow=True)[0]`` (``psld.py:345-364``) with ``u`` in its own
    dtype: python scalars times a float32 state stay float32 inside the drift bracket (``psld.py:336-337``),
    the float64 ``beta_t`` promotes afterwards; the fp32 score is halved in fp32."""
    tau = sde.T - t
    x, m = torch.chunk(u, 2, dim=1)
    beta = sde.beta_t(tau)
    fx = (0.5 * beta) * (sde.m_inv * m - sde.gamma * x).to(torch.float64)
    fm = (0.5 * beta) * (-sde.nu * m - x).to(torch.float64)
    gx2, gm2 = math.sqrt(beta * sde.gamma) ** 2, math.sqrt(beta * sde.m * sde.nu) ** 2
    eps = score_fn(u.to(torch.float32), _tvec(u, float(np.float32(tau))))
    sx, sm = torch.chunk(0.5 * score_from_eps(sde, eps, tau), 2, dim=1)
    return torch.cat([-fx + gx2 * sx, -fm + gm2 * sm], dim=1)


def bb_ode_sample(config, score_fn, u0, rtol, atol, solver="RK45", denoise=True, eps=1e-3):
    """``BBODESampler.sample`` (``ode.py:41-76``): probability-flow ODE integrated from t = 0 to T - eps
    by ``torchdiffeq.odeint(method="scipy_solver")``, restated as the ``scipy.integrate.solve_ivp`` call
    that wrapper makes (torchdiffeq 0.2.3 is absent from the reference tree): float64 state on the
    host, ``(t, y)`` rounded to ``u0``'s dtype for every evaluation.  Returns ``(x, nfe)``."""
    from scipy.integrate import solve_ivp
    vp = str(config.model.sde.name) == "vpsde"        # sample_uncond_vpsde_ode.sh: same sampler, VP-SDE
    sde = VPScalars(config) if vp else PSLDScalars(config)
    drift = vp_pflow_drift if vp else pflow_drift
    shape, dtype = u0.shape, u0.dtype
    nfe = [0]

    def fun(t, y):
        nfe[0] += 1
        tt = float(torch.tensor(t).to(dtype))
        u = torch.tensor(y).to(dtype).reshape(shape)
        with torch.no_grad():
            return drift(sde, score_fn, u, tt).numpy().reshape(-1)

    t_end = sde.T - eps
    sol = solve_ivp(fun, t_span=[0.0, t_end], y0=u0.numpy().reshape(-1), t_eval=np.asarray([0.0, t_end]),
                    method=solver, rtol=rtol, atol=atol, max_step=float("inf"))
    x = torch.tensor(sol.y).T[-1].to(dtype).reshape(shape)
    if denoise:                                                    # ode.py:36-39,66-75
        with torch.no_grad():
            x = x + drift(sde, score_fn, x, sde.T - eps) * eps
        nfe[0] += 1
    return x, nfe[0]


def sscs_sample(config, score_fn, u0, ts, n, noise, denoise=True, eps=1e-3, record=None):
    """``SSCSSampler.sample`` (``sde.py:350-370``).  ``noise``: 2n tensors in draw order
    z1(step0), z2(step0), z1(step1), ...  (the denoise step's extra draw is discarded by
    the reference, ``sde.py:346``, and is not needed here)."""
    sde = PSLDScalars(config)
    u = u0.to(torch.float64)
    with torch.no_grad():
        for i in range(n):
            t, dt = float(ts[i]), float(ts[i + 1] - ts[i])
            # both half-steps use the same t (reference behaviour, sde.py:333-335)
            (axx, axm, amx, amm), (c11, c12, c21, c22) = sde.half_step(t, dt / 2)

            def half(u, z):
                x, m = torch.chunk(u, 2, dim=1)
                zx, zm = torch.chunk(z.to(torch.float64), 2, dim=1)
                return torch.cat([axx * x + axm * m + (c11 * zx + c12 * zm),
                                  amx * x + amm * m + (c21 * zx + c22 * zm)], dim=1)

            u = half(u, noise[2 * i])
            # score step (sde.py:314-329)
            tau = sde.T - t
            beta = sde.beta_t(tau)
            e = score_fn(u.to(torch.float32), _tvec(u, float(np.float32(tau))))
            score = score_from_eps(sde, e, tau)
            x, m = torch.chunk(u, 2, dim=1)
            sx, sm = torch.chunk(score, 2, dim=1)
            x = x + dt * sde.gamma * beta * (sx + x)
            m = m + dt * sde.m * sde.nu * beta * (sm + sde.m_inv * m)
            u = half(torch.cat([x, m], dim=1), noise[2 * i + 1])
            if record is not None:
                record(i, u)
        if denoise:                                                # sde.py:338-348,364-369
            u = _denoise(sde, score_fn, u, eps)
    return u


def mean_coeffs(sde: PSLDScalars, tau: float):
    """``PSLD._mean`` (``psld.py:62-84``) as coefficients: mu_x = a_xx x0 + a_xm m0, mu_m = ..."""
    mu_lam = (sde.nu + sde.gamma) / 4
    b = sde.b_t(tau)
    s = math.exp(-mu_lam * b)
    a1, a2 = (sde.nu - sde.gamma) / 4, (sde.gamma - sde.nu) ** 2 / 8
    c1, c2 = -0.5, (sde.gamma - sde.nu) / 4
    return s * (1 + a1 * b), s * (a2 * b), s * (c1 * b), s * (1 + c2 * b)


def inpaint_em_sample(config, score_fn, x_0, mask, prior, ts, n, noise, denoise=True, eps=1e-3,
                      record=None):
    """``ES3EulerMaruyamaInpainter.sample`` (``sde.py:188-224``).  ``prior`` = the tensor
    ``sde.prior_sampling`` returned; ``noise`` = dict of draws in the reference's order:
    ``pred[i]`` (predictor, sde.py:158), ``m0[k]`` / ``eps[k]`` (``_perturb``, sde.py:137-146) with
    k = 0 for the initial latent, i+1 after step i and n+1 for the denoise call."""
    sde = PSLDScalars(config)
    hsm = str(config.training.mode) == "hsm"
    C = x_0.shape[1]
    x0 = x_0.to(torch.float64)
    mk = mask.to(torch.float64)

    def perturb(tau, k):                                           # sde.py:134-150
        m_0 = math.sqrt(sde.mm_0) * noise["m0"][k].to(torch.float64)
        mm_0 = 0.0
        if hsm:
            m_0 = torch.zeros_like(x0)
            mm_0 = sde.mm_0
        axx, axm, amx, amm = mean_coeffs(sde, tau)
        mu = torch.cat([axx * x0 + axm * m_0, amx * x0 + amm * m_0], dim=1)
        c11, c12, c21, c22 = sde.get_coeff(sde.cov(0.0, mm_0, tau))
        ex, em = torch.chunk(noise["eps"][k].to(torch.float64), 2, dim=1)
        return mu + torch.cat([c11 * ex + c12 * em, c21 * ex + c22 * em], dim=1), mu

    def combine(u, other):                                         # sde.py:174-178
        a, b = torch.chunk(u, 2, dim=1)
        ok, om = torch.chunk(other, 2, dim=1)
        return torch.cat([a * (1 - mk) + ok * mk, b * (1 - mk) + om * mk], dim=1)

    u = combine(prior.to(torch.float64), perturb(sde.T, 0)[0])     # sde.py:193-202
    with torch.no_grad():
        for i in range(n):
            t, dt = float(ts[i]), float(ts[i + 1] - ts[i])
            fbar, (gx, gm) = reverse_drift(sde, score_fn, u, t)
            g = torch.cat([torch.full_like(u[:, :C], gx), torch.full_like(u[:, C:], gm)], dim=1)
            u = (u + fbar * dt) + g * math.sqrt(dt) * noise["pred"][i].to(torch.float64)
            u = combine(u, perturb(sde.T - t, i + 1)[0])           # perturbed at T - ts[i], sde.py:168-172
            if record is not None:
                record(i, u)
        if denoise:                                                # sde.py:213-222: float32 t and dt
            t32 = np.float32(sde.T - eps)
            fbar, _ = reverse_drift(sde, score_fn, u, float(t32))
            mean = u + fbar * float(np.float32(eps))
            u = combine(mean, perturb(float(np.float32(sde.T) - t32), n + 1)[1])
    return u


class VPScalars:
    """Scalar restatement of ``VPSDE`` (``vpsde.py:8-99``)."""

    def __init__(self, config):
        self.beta_0, self.beta_1 = float(config.model.sde.beta_min), float(config.model.sde.beta_max)
        self.T = 1.0

    def beta_t(self, t):                                           # vpsde.py:17-18
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def std(self, t):                                              # vpsde.py:88-92
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return math.sqrt(1.0 - math.exp(2.0 * lmc))


def vp_reverse_drift(sde: VPScalars, score_fn, x: torch.Tensor, t: float):
    """``VPSDE.reverse_sde`` (``vpsde.py:51-74``), probability_flow=False: (f_bar, g)."""
    tau = sde.T - t
    beta = sde.beta_t(tau)
    g = math.sqrt(beta)
    eps = score_fn(x.to(torch.float32), _tvec(x, float(np.float32(tau))))
    score = -eps.to(torch.float64) / sde.std(tau)                  # vpsde.py:27-28
    return 0.5 * beta * x + g ** 2 * score, g


def vp_pflow_drift(sde: VPScalars, score_fn, x: torch.Tensor, t: float):
    """``VPSDE.reverse_sde(x, t, score_fn, probability_flow=True)[0]`` (``vpsde.py:48-67``): beta_t and std
    are float64 tensors in the reference, so a float32 state / eps promote before the first product; the
    float64 score is halved."""
    tau = sde.T - t
    beta = sde.beta_t(tau)
    g = math.sqrt(beta)
    eps = score_fn(x.to(torch.float32), _tvec(x, float(np.float32(tau))))
    score = 0.5 * (-eps.to(torch.float64) / sde.std(tau))
    f = (-0.5 * beta) * x.to(torch.float64)
    return -f + g ** 2 * score


def vp_em_sample(config, score_fn, x0, ts, n, noise, denoise=True, eps=1e-3, record=None):
    """``EulerMaruyamaSampler.sample`` (``sde.py:38-58``) on the VP-SDE; state ``[B,C,H,W]``."""
    sde = VPScalars(config)
    x = x0.to(torch.float64)
    with torch.no_grad():
        for i in range(n):
            t, dt = float(ts[i]), float(ts[i + 1] - ts[i])
            fbar, g = vp_reverse_drift(sde, score_fn, x, t)
            x = (x + fbar * dt) + g * math.sqrt(dt) * noise[i].to(torch.float64)
            if record is not None:
                record(i, x)
        if denoise:                                                # float32 t, dt (sde.py:52-57)
            fbar, _ = vp_reverse_drift(sde, score_fn, x, float(np.float32(sde.T - eps)))
            x = x + fbar * float(np.float32(eps))
    return x


# --------------------------------------------------------------------------------------
# 3. upfirdn2d (op/upfirdn2d.py:159-200): zero-stuff, pad/crop, TRUE convolution, decimate
# --------------------------------------------------------------------------------------
def upfirdn2d(x: torch.Tensor, k, up=1, down=1, pad=(0, 0)) -> torch.Tensor:
    """x: [N,C,H,W]; k: [kh,kw] array-like; same ``pad`` on both axes as the reference's
    public wrapper (``op/upfirdn2d.py:145-156``)."""
    k = torch.as_tensor(np.asarray(k, dtype=np.float32))
    N, C, H, W = x.shape
    kh, kw = k.shape
    p0, p1 = pad
    z = x.new_zeros(N, C, H * up, W * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    z = z[:, :, max(-p0, 0): z.shape[2] - max(-p1, 0), max(-p0, 0): z.shape[3] - max(-p1, 0)]
    Hp, Wp = z.shape[2], z.shape[3]
    w = torch.flip(k, [0, 1]).view(1, 1, kh, kw).to(device=x.device, dtype=x.dtype)
    o = F.conv2d(z.reshape(N * C, 1, Hp, Wp), w).reshape(N, C, Hp - kh + 1, Wp - kw + 1)
    return o[:, :, ::down, ::down]


def fir_kernel_2d(k1d, gain=1.0):
    """``up_or_down_sampling.py:181-188``."""
    k = np.asarray(k1d, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k = k / np.sum(k)
    return k * gain


def upsample_2d(x, k1d, factor=2):                                 # :195-224
    k = fir_kernel_2d(k1d, factor ** 2)
    p = k.shape[0] - factor
    return upfirdn2d(x, k, up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k1d, factor=2):                               # :227-257
    k = fir_kernel_2d(k1d)
    p = k.shape[0] - factor
    return upfirdn2d(x, k, down=factor, pad=((p + 1) // 2, p // 2))


def conv_downsample_2d(x, w, k1d, factor=2):                       # :144-178
    k = fir_kernel_2d(k1d)
    p = (k.shape[0] - factor) + (w.shape[-1] - 1)
    x = upfirdn2d(x, k, pad=((p + 1) // 2, p // 2))
    return F.conv2d(x, w, stride=factor, padding=0)


# --------------------------------------------------------------------------------------
# 4. NCSN++ forward (functional, from a state dict)
# --------------------------------------------------------------------------------------
def _gn(sd, key, x):
    C = x.shape[1]
    return F.group_norm(x, min(C // 4, 32), sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _nin(sd, key, x):                                              # layers.py:531-540
    y = torch.einsum("bhwc,cd->bhwd", x.permute(0, 2, 3, 1), sd[key + ".W"]) + sd[key + ".b"]
    return y.permute(0, 3, 1, 2)


def resblock(sd, p, x, temb, cfg_sf, up=False, down=False):       # layerspp.py:242-274
    fk = cfg_sf.fir_kernel
    in_ch = x.shape[1]
    out_ch = sd[p + ".Conv_0.weight"].shape[0]
    h = F.silu(_gn(sd, p + ".GroupNorm_0", x))
    if up:
        if cfg_sf.fir:
            h, x = upsample_2d(h, fk), upsample_2d(x, fk)
        else:
            h, x = [F.interpolate(v, scale_factor=2, mode="nearest") for v in (h, x)]
    elif down:
        if cfg_sf.fir:
            h, x = downsample_2d(h, fk), downsample_2d(x, fk)
        else:
            h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    h = F.conv2d(h, sd[p + ".Conv_0.weight"], sd[p + ".Conv_0.bias"], padding=1)
    if temb is not None:
        h = h + F.linear(F.silu(temb), sd[p + ".Dense_0.weight"], sd[p + ".Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(sd, p + ".GroupNorm_1", h))
    h = F.conv2d(h, sd[p + ".Conv_1.weight"], sd[p + ".Conv_1.bias"], padding=1)
    if in_ch != out_ch or up or down:
        x = F.conv2d(x, sd[p + ".Conv_2.weight"], sd[p + ".Conv_2.bias"])
    return (x + h) / np.sqrt(2.0) if cfg_sf.skip_rescale else x + h


def attnblock(sd, p, x, skip_rescale=True):                        # layerspp.py:75-91
    B, C, H, W = x.shape
    h = _gn(sd, p + ".GroupNorm_0", x)
    q, k, v = (_nin(sd, p + f".NIN_{i}", h) for i in range(3))
    w = torch.einsum("bchw,bcij->bhwij", q, k) * (int(C) ** (-0.5))
    w = F.softmax(w.reshape(B, H, W, H * W), dim=-1).reshape(B, H, W, H, W)
    h = torch.einsum("bhwij,bcij->bchw", w, v)
    h = _nin(sd, p + ".NIN_3", h)
    return (x + h) / np.sqrt(2.0) if skip_rescale else x + h


def ncsnpp_forward(config, sd, x, time_cond):
    """``NCSNpp.forward`` for resblock_type=biggan, progressive=none,
    progressive_input in {none,residual}, embedding_type in {fourier,positional}
    (``ncsnpp.py:287-438``).  ``sd`` maps ``all_modules.<i>.<...>`` → fp32 tensors."""
    sf = config.model.score_fn
    nf, ch_mult, nrb = sf.nf, list(sf.ch_mult), sf.num_res_blocks
    nres = len(ch_mult)
    resolutions = [config.data.image_size // (2 ** i) for i in range(nres)]
    attn_res = list(sf.attn_resolutions)
    assert sf.resblock_type.lower() == "biggan" and sf.progressive.lower() == "none"
    pin = sf.progressive_input.lower()
    assert pin in ("none", "residual")
    key = lambda i: f"all_modules.{i}"
    i = 0
    if sf.embedding_type.lower() == "fourier":                     # ncsnpp.py:292-296
        W = sd[key(i) + ".W"]; i += 1
        xp = torch.log(time_cond)[:, None] * W[None, :] * 2 * np.pi    # layerspp.py:40
        temb = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
    else:                                                          # layers.py:500-514
        half = nf // 2
        e = math.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
        e = time_cond.float()[:, None] * e[None, :]
        temb = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if sf.noise_cond:                                              # ncsnpp.py:307-311
        temb = F.linear(temb, sd[key(i) + ".weight"], sd[key(i) + ".bias"]); i += 1
        temb = F.linear(F.silu(temb), sd[key(i) + ".weight"], sd[key(i) + ".bias"]); i += 1
    else:
        temb = None
    pyr = x if pin != "none" else None
    hs = [F.conv2d(x, sd[key(i) + ".weight"], sd[key(i) + ".bias"], padding=1)]; i += 1
    for lvl in range(nres):
        for _ in range(nrb):
            h = resblock(sd, key(i), hs[-1], temb, sf); i += 1
            if h.shape[-1] in attn_res:
                h = attnblock(sd, key(i), h, sf.skip_rescale); i += 1
            hs.append(h)
        if lvl != nres - 1:
            h = resblock(sd, key(i), hs[-1], temb, sf, down=True); i += 1
            if pin == "residual":                                  # ncsnpp.py:350-357
                p = key(i) + ".Conv2d_0"; i += 1
                if sf.fir:
                    pyr = conv_downsample_2d(pyr, sd[p + ".weight"], sf.fir_kernel) \
                        + sd[p + ".bias"].reshape(1, -1, 1, 1)
                else:
                    raise NotImplementedError("progressive_input=residual requires fir")
                pyr = (pyr + h) / np.sqrt(2.0) if sf.skip_rescale else pyr + h
                h = pyr
            hs.append(h)
    h = hs[-1]
    h = resblock(sd, key(i), h, temb, sf); i += 1
    h = attnblock(sd, key(i), h, sf.skip_rescale); i += 1
    h = resblock(sd, key(i), h, temb, sf); i += 1
    for lvl in reversed(range(nres)):
        for _ in range(nrb + 1):
            h = resblock(sd, key(i), torch.cat([h, hs.pop()], dim=1), temb, sf); i += 1
        if h.shape[-1] in attn_res:
            h = attnblock(sd, key(i), h, sf.skip_rescale); i += 1
        if lvl != 0:
            h = resblock(sd, key(i), h, temb, sf, up=True); i += 1
    assert not hs
    h = F.silu(_gn(sd, key(i), h)); i += 1
    h = F.conv2d(h, sd[key(i) + ".weight"], sd[key(i) + ".bias"], padding=1); i += 1
    n_mod = 1 + max(int(k.split(".")[1]) for k in sd)
    assert i == n_mod, (i, n_mod)
    return h


class OracleScoreFn:
    """score_fn(u f32 [B,in_ch,H,W], t f32 [B]) -> eps f32, backed by ``ncsnpp_forward``."""

    def __init__(self, config, sd):
        self.config = config
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}

    def __call__(self, u, t):
        with torch.no_grad():
            return ncsnpp_forward(self.config, self.sd, u, t)


class GuidedScoreFn:
    """Classifier-free guidance over two score functions: eps = (1 + w) eps_c - w eps_u, evaluated in
    float32 in exactly this order (two products, one sum).  BASELINE configs[4]; the reference ships
    no such sampler (its NCSN++ takes no label, ncsnpp.py:288), so this composition of two reference
    forwards at the reference's score_fn call site (sde.py:320, psld.py:354) is the definition the
    CUDA path is held to - parity is pinned on the two forwards, not on a reference CFG run."""

    def __init__(self, cond, uncond, weight):
        self.cond, self.uncond, self.weight = cond, uncond, float(weight)

    def __call__(self, u, t):
        a = torch.tensor(1.0 + self.weight, dtype=torch.float32, device=u.device)
        b = torch.tensor(-self.weight, dtype=torch.float32, device=u.device)
        return a * self.cond(u, t) + b * self.uncond(u, t)


# --------------------------------------------------------------------------------------
# 5. Caller-side image quantisation (callbacks.py:103-107, util.py:147-158)
# --------------------------------------------------------------------------------------
def images_uint8(state: torch.Tensor) -> np.ndarray:
    """[B,2C,H,W] -> uint8 [B,H,W,C]: drop momentum, x*0.5+0.5, *255, clip, astype(uint8)."""
    x, _ = torch.chunk(state.cpu(), 2, dim=1)
    obj = x * 0.5 + 0.5
    arr = obj.permute(0, 2, 3, 1).contiguous().numpy()
    return (arr * 255).clip(0, 255).astype(np.uint8)
