"""Host-side PSLD schedule: every per-step scalar of the sampler, once, in float64.

The reference recomputes these as ``(B,)`` float64 device tensors on every step —
identical across the batch — with ~600 tiny launches and 9 host syncs per SSCS step
(SURVEY.md §3.1).  Here they are built once per ``sample()`` call as tables of C
structs (``include/psld_b200.h``) that the fused kernels take by value.

Reference lines restated (relative to the reference repo):
  beta_t / b_t ................ main/models/sde/psld.py:38-44
  _cov (perturbation kernel) .. main/models/sde/psld.py:86-152
  get_coeff / get_inv_coeff ... main/models/sde/psld.py:154-220 (NaN -> ValueError :171)
  SSCS _mean / _var ........... main/samplers/sde.py:236-292
  euler_score_dynamics ........ main/samplers/sde.py:314-329
  sde / reverse_sde ........... main/models/sde/psld.py:330-364
  time grid ................... main/models/wrapper.py:51-54,101-114

All transcendental functions go through torch float64 CPU ops (the same library calls
the reference makes), so the coefficients agree with the reference's to the last bit
or two.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L

_F64 = torch.float64


class PSLDSchedule:
    """SDE constants (reference ``PSLD.__init__``, psld.py:14-33)."""

    def __init__(self, config):
        c = config.model.sde
        self.beta_0 = float(c.beta_min)
        self.beta_1 = float(c.beta_max)
        self.nu = float(c.nu)
        self.gamma = float(c.gamma)
        assert self.nu != 0 or self.gamma != 0
        self.m_inv = (self.gamma - self.nu) ** 2 / 4
        self.m = 1 / self.m_inv
        self.kappa = float(c.kappa)
        self.mm_0 = self.kappa * self.m
        self.eps = float(c.numerical_eps)
        self.decomp_mode = str(c.decomp_mode)
        assert self.decomp_mode in ["lower", "upper"]
        self.T = 1.0

    @classmethod
    def from_sde(cls, sde):
        """Builds from a reference ``PSLD`` instance (duck-typed attributes)."""
        self = cls.__new__(cls)
        for k in ("beta_0", "beta_1", "nu", "gamma", "m_inv", "m", "kappa", "mm_0", "eps",
                  "decomp_mode"):
            setattr(self, k, getattr(sde, k))
        self.T = float(getattr(sde, "T", 1.0))
        return self

    # ---- mode (psld.py:50-56) ----
    @property
    def mode(self):
        if self.gamma == 0:
            return "score_m"
        if self.nu == 0:
            return "score_x"
        return "score_xm"

    @property
    def mode_code(self):
        """0: eps has 2C channels; 1: score_m (lower) ; 2: score_x (upper) — psld.py:240-248."""
        if self.decomp_mode == "lower" and self.mode == "score_m":
            return 1
        if self.decomp_mode == "upper" and self.mode == "score_x":
            return 2
        return 0

    def beta_t(self, t):
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def b_t(self, t):
        return self.beta_0 * t + 0.5 * (t ** 2) * (self.beta_1 - self.beta_0)

    # ---- perturbation-kernel covariance at forward time tau, from (xx0, mm0) ----
    def cov(self, xx_0, mm_0, tau):
        nu, ga, mi, m = self.nu, self.gamma, self.m_inv, self.m
        lam = (nu + ga) / 2
        b = self.b_t(tau)
        b2 = b ** 2
        s = torch.exp(-lam * b)
        si = torch.exp(lam * b)
        xx = (mi / 4 * b2 * xx_0 + mi ** 2 / 4 * b2 * mm_0 + (nu - ga) / 2 * b * xx_0
              + (-mi / 2) * b2 + (ga - nu) / 2 * b + (si - 1) + xx_0) * s
        xm = ((ga - nu) / 8 * b2 * xx_0 + mi * (ga - nu) / 8 * b2 * mm_0 + (-1 / 2) * b * xx_0
              + mi / 2 * b * mm_0 + (nu - ga) / 4 * b2) * s
        mm = (1 / 4 * b2 * xx_0 + mi / 4 * b2 * mm_0 + (ga - nu) / 2 * b * mm_0
              + (-1 / 2) * b2 + m * (nu - ga) / 2 * b + m * (si - 1) + mm_0) * s
        return xx + self.eps, xm, mm + self.eps

    def factor(self, var):
        """(c11, c12, c21, c22) with C C^T = var; lower Cholesky or the upper variant."""
        xx, xm, mm = var
        zero = torch.zeros_like(xx)
        if self.decomp_mode == "lower":
            l11 = torch.sqrt(xx)
            l21 = xm / l11
            l22 = torch.sqrt(mm - l21 ** 2.0)
            out = (l11, zero, l21, l22)
        else:
            u22 = torch.sqrt(mm)
            u12 = xm / u22
            u11 = torch.sqrt(xx - u12 ** 2.0)
            out = (u11, u12, zero, u22)
        _nan_guard(out)
        return out

    def inv_factor_T(self, var):
        """Entries of the inverse-transpose of :meth:`factor` (psld.py:188-220)."""
        xx, xm, mm = var
        det = xx * mm - xm ** 2
        zero = torch.zeros_like(xx)
        if self.decomp_mode == "lower":
            out = (torch.sqrt(1 / xx), -xm / (torch.sqrt(xx) * torch.sqrt(det)), zero,
                   torch.sqrt(xx / det))
        else:
            out = (torch.sqrt(mm / det), zero, -xm / (torch.sqrt(mm) * torch.sqrt(det)),
                   torch.sqrt(1 / mm))
        _nan_guard(out)
        return out

    # ---- mean of the perturbation kernel as a 2x2 map of (x_0, m_0) (psld.py:62-84) ----
    def mean_coeffs(self, tau):
        mu_lam = (self.nu + self.gamma) / 4
        b = self.b_t(tau)
        s = torch.exp(-mu_lam * b)
        a1 = (self.nu - self.gamma) / 4
        a2 = (self.gamma - self.nu) ** 2 / 8
        c2 = (self.gamma - self.nu) / 4
        return s * (1 + a1 * b), s * (a2 * b), s * (-0.5 * b), s * (1 + c2 * b)

    # ---- SSCS analytic half-step over [t, t+h] in reverse time (sde.py:236-292) ----
    def half_step(self, t, h):
        nu, ga = self.nu, self.gamma
        db = self.b_t(self.T - (t + h)) - self.b_t(self.T - t)
        s = torch.exp((nu + ga) / 4 * db)
        a_xx = s * (1 - (nu - ga) / 4 * db)
        a_xm = s * ((ga - nu) ** 2 / 8 * db)
        a_mx = s * (-0.5 * db)
        a_mm = s * (1 - (ga - nu) / 4 * db)
        lam = (nu + ga) / 2
        E, Ei = torch.exp(lam * db), torch.exp(-lam * db)
        db2 = db ** 2
        xx = (-self.m_inv / 2 * db2 - (ga - nu) / 2 * db + (Ei - 1)) * E + self.eps
        xm = ((ga - nu) / 4 * db2) * E
        mm = (-0.5 * db2 - self.m * (nu - ga) / 2 * db + self.m * (Ei - 1)) * E + self.eps
        return (a_xx, a_xm, a_mx, a_mm), self.factor((xx, xm, mm))


def _nan_guard(vals):
    for v in vals:
        if bool(torch.isnan(v).any()):
            raise ValueError("Numerical precision error.")      # same text as psld.py:171


def time_grid(config, T=1.0, device="cpu"):
    """``ts`` exactly as the caller builds it (wrapper.py:51-54,101-114) -> (ts f64, n)."""
    ev = config.evaluation
    n = ev.n_discrete_steps - 1 if ev.denoise else ev.n_discrete_steps
    t_final = T - ev.eval_eps
    ts = torch.linspace(0, t_final, n + 1, dtype=_F64, device=device)
    if ev.stride_type == "quadratic":
        ts = t_final * torch.flip(1 - (ts / t_final) ** 2.0, dims=[0])
    return ts, n


def _fill_half(dst: L.HalfStep, a, c, i):
    dst.a_xx, dst.a_xm, dst.a_mx, dst.a_mm = (float(v[i]) for v in a)
    dst.c11, dst.c12, dst.c21, dst.c22 = (float(v[i]) for v in c)


def _score_rows(sch: PSLDSchedule, tau, dt, sqrt_dt=None):
    """Per-step score/drift scalars at forward times ``tau`` (f64 [n]) with steps ``dt``."""
    li = sch.inv_factor_T(sch.cov(0.0, sch.mm_0, tau))
    li32 = [v.to(torch.float32) for v in li]                  # psld.py:252-259 `.type(float32)`
    beta = sch.beta_t(tau)
    g_x = torch.sqrt(beta * sch.gamma)                        # psld.py:339
    g_m = torch.sqrt(beta * sch.m * sch.nu)                   # psld.py:340
    rows = dict(
        li=li32, beta=beta,
        k_x=dt * sch.gamma * beta,                            # sde.py:325
        k_m=dt * sch.m * sch.nu * beta,                       # sde.py:326
        g2_x=g_x ** 2, g2_m=g_m ** 2,
        gs_x=g_x * sqrt_dt if sqrt_dt is not None else torch.zeros_like(beta),
        gs_m=g_m * sqrt_dt if sqrt_dt is not None else torch.zeros_like(beta),
        dt=dt,
    )
    return rows


def _fill_score(dst: L.ScoreStep, sch: PSLDSchedule, rows, i):
    dst.li11, dst.li12, dst.li21, dst.li22 = (float(v[i]) for v in rows["li"])
    dst.mode = sch.mode_code
    dst.k_x = float(rows["k_x"][i])
    dst.k_m = float(rows["k_m"][i])
    dst.m_inv = sch.m_inv
    dst.half_beta = float(0.5 * rows["beta"][i])
    dst.gamma, dst.nu = sch.gamma, sch.nu
    dst.g2_x = float(rows["g2_x"][i])
    dst.g2_m = float(rows["g2_m"][i])
    dst.dt = float(rows["dt"][i])
    dst.gs_x = float(rows["gs_x"][i])
    dst.gs_m = float(rows["gs_m"][i])


def _merge_halves(hb: L.HalfStep, hc: L.HalfStep, dst: L.HalfStep):
    """One Gaussian draw for two consecutive half-steps: u'' = A_c (A_b u + L_b z_b) + L_c z_c has
    mean A_c A_b u and covariance A_c L_b L_b^T A_c^T + L_c L_c^T, so it equals (in law, exactly)
    A_bc u + L_bc z with L_bc the lower Cholesky factor of that covariance.  Used only when the
    noise comes from the in-kernel generator (nothing to match draw by draw)."""
    import numpy as np
    Ab = np.array([[hb.a_xx, hb.a_xm], [hb.a_mx, hb.a_mm]])
    Ac = np.array([[hc.a_xx, hc.a_xm], [hc.a_mx, hc.a_mm]])
    Lb = np.array([[hb.c11, hb.c12], [hb.c21, hb.c22]])
    Lc = np.array([[hc.c11, hc.c12], [hc.c21, hc.c22]])
    A = Ac @ Ab
    S = Ac @ Lb @ Lb.T @ Ac.T + Lc @ Lc.T
    l11 = math.sqrt(S[0, 0]); l21 = S[1, 0] / l11; l22 = math.sqrt(S[1, 1] - l21 * l21)
    if any(math.isnan(v) for v in (l11, l21, l22)):
        raise ValueError("Numerical precision error.")
    dst.a_xx, dst.a_xm, dst.a_mx, dst.a_mm = A[0, 0], A[0, 1], A[1, 0], A[1, 1]
    dst.c11, dst.c12, dst.c21, dst.c22 = l11, 0.0, l21, l22


class StepTables:
    """ctypes tables for one ``sample()`` call."""

    def __init__(self, sch: PSLDSchedule, ts, n: int, sampler: str, denoise: bool, eps: float,
                 embedding: str = "fourier", merge_noise: bool = False, dt=None):
        """``dt`` (optional f64 [n]) overrides ``ts[i+1] - ts[i]``: a standalone
        ``predictor_update_fn(u, t, dt)`` call passes its step verbatim."""
        ts = torch.as_tensor(ts, dtype=_F64).cpu()
        assert ts.numel() >= n + 1 or (dt is not None and ts.numel() >= n)
        self.n = n
        t = ts[:n]
        dt = ts[1:n + 1] - ts[:n] if dt is None else torch.as_tensor(dt, dtype=_F64).cpu().reshape(-1)[:n]
        tau = sch.T - t
        self.sscs = self.em = None
        if n > 0:
            if sampler == "sscs_sde":
                a, c = sch.half_step(t, dt / 2)               # same t for both halves (sde.py:333-335)
                rows = _score_rows(sch, tau, dt)
                tab = (L.SscsCoeffs * n)()
                for i in range(n):
                    _fill_half(tab[i].half_a, a, c, i)
                    _fill_half(tab[i].half_b, a, c, i)
                    if i + 1 < n:
                        _fill_half(tab[i].half_c, a, c, i + 1)
                    _fill_score(tab[i].score, sch, rows, i)
                if merge_noise:      # half B of step i and half A of step i+1 share one draw
                    for i in range(n - 1):
                        _merge_halves(tab[i].half_b, tab[i].half_c, tab[i].half_b)
                self.sscs = tab
            elif sampler == "em_sde":
                rows = _score_rows(sch, tau, dt, torch.sqrt(dt))  # sde.py:24 `g * sqrt(dt)`
                tab = (L.ScoreStep * n)()
                for i in range(n):
                    _fill_score(tab[i], sch, rows, i)
                self.em = tab
            else:
                raise ValueError(f"unknown sampler {sampler}")
        # denoise: x + fbar * eps at t = T - eps, i.e. tau = eps (sde.py:28-36,338-348)
        self.den = None
        call_tau = [tau] if n > 0 else []
        if denoise:
            # the reference passes t = torch.tensor(T - eps) and dt = torch.tensor(eps), both
            # float32 tensors (sde.py:52-57, 364-369): keep that rounding
            tden = torch.tensor([sch.T - eps], dtype=torch.float32).to(_F64)
            tau_d = sch.T - tden
            rows = _score_rows(sch, tau_d, torch.tensor([eps], dtype=torch.float32).to(_F64))
            self.den = L.ScoreStep()
            _fill_score(self.den, sch, rows, 0)
            call_tau.append(tau_d)
        # network time per call, as the reference feeds it: tau -> float32 (sde.py:320), then
        # log in float32 for the Fourier embedding (ncsnpp.py:295)
        tau_all = torch.cat(call_tau) if call_tau else torch.zeros(0, dtype=_F64)
        self.tau32 = tau_all.to(torch.float32)
        self.time_table = torch.log(self.tau32) if embedding == "fourier" else self.tau32.clone()

    @property
    def n_calls(self):
        return int(self.tau32.numel())


class InpaintTables:
    """Per-call coefficients of the inpainting sampler's Split-Perturb-Combine
    (reference sde.py:134-186): call 0 = initial latent (t = T), calls 1..n = after predictor
    step i at forward time T - ts[i] (the reference perturbs at the time the step STARTED from,
    sde.py:168-172), call n+1 = the denoise call at T - fl32(T - eps), mean only."""

    def __init__(self, sch: PSLDSchedule, ts, n: int, denoise: bool, eps: float, hsm: bool):
        ts = torch.as_tensor(ts, dtype=_F64).cpu()
        taus = [torch.tensor([sch.T], dtype=_F64), sch.T - ts[:n]]
        if denoise:
            # `self.sde.T - t` with t = torch.tensor(T - eps) is a float32 subtraction (sde.py:211-219)
            t32 = torch.tensor(sch.T - eps, dtype=torch.float32)
            taus.append((sch.T - t32).to(_F64).reshape(1))
        tau = torch.cat(taus)
        mm_0 = sch.mm_0 if hsm else 0.0
        a = sch.mean_coeffs(tau)
        c = sch.factor(sch.cov(0.0, mm_0, tau))
        k = int(tau.numel())
        self.steps = (L.InpaintStep * k)()
        for i in range(k):
            st = self.steps[i]
            st.a_xx, st.a_xm, st.a_mx, st.a_mm = (float(v[i]) for v in a)
            st.c11, st.c12, st.c21, st.c22 = (float(v[i]) for v in c)
            st.m0_std = 0.0 if hsm else float(sch.mm_0) ** 0.5
            st.mean_only = 0
        if denoise:
            self.steps[k - 1].mean_only = 1
        self.tau = tau


class VPSchedule:
    """Scalars of the VP-SDE baseline (reference ``VPSDE``, vpsde.py:8-99)."""

    def __init__(self, config):
        self.beta_0 = float(config.model.sde.beta_min)
        self.beta_1 = float(config.model.sde.beta_max)

    @classmethod
    def from_sde(cls, sde):
        self = cls.__new__(cls)
        self.beta_0, self.beta_1 = float(sde.beta_0), float(sde.beta_1)
        return self

    T = 1.0

    def beta_t(self, t):
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def log_mean_coeff(self, t):                                   # vpsde.py:78-80
        return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

    def std(self, t):                                              # vpsde.py:88-92
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_mean_coeff(t)))


class VPStepTables:
    """Per-step coefficients of Euler-Maruyama on the VP-SDE (sde.py:16-36 x vpsde.py:42-74)."""

    def __init__(self, sch: VPSchedule, ts, n: int, denoise: bool, eps: float, embedding="fourier"):
        ts = torch.as_tensor(ts, dtype=_F64).cpu()
        t, dt = ts[:n], ts[1:n + 1] - ts[:n]
        taus, dts = [sch.T - t], [dt]
        if denoise:        # float32 t and dt of the denoising call (sde.py:52-57)
            taus.append(sch.T - torch.tensor([sch.T - eps], dtype=torch.float32).to(_F64))
            dts.append(torch.tensor([eps], dtype=torch.float32).to(_F64))
        tau, dta = torch.cat(taus), torch.cat(dts)
        beta = sch.beta_t(tau)
        g = torch.sqrt(beta)
        k = int(tau.numel())
        self.steps = (L.VpStep * k)()
        for i in range(k):
            st = self.steps[i]
            st.half_beta = float(0.5 * beta[i])
            st.g2 = float(g[i] ** 2)
            st.neg_inv_std = float(-1.0 / sch.std(tau[i]))
            st.dt = float(dta[i])
            st.gs = float(g[i] * torch.sqrt(dta[i])) if i < n else 0.0
        self.tau32 = tau.to(torch.float32)
        self.time_table = torch.log(self.tau32) if embedding == "fourier" else self.tau32.clone()
