"""PSLD SDE object with the reference's attribute surface for the sampling path.

Mirrors ``PSLD`` (reference ``main/models/sde/psld.py:12-60,366-370``): same constructor
(``PSLD(config)``), same attributes (``beta_0, beta_1, nu, gamma, m_inv, m, kappa, mm_0, eps,
decomp_mode, T, mode``), ``beta_t``/``b_t`` and ``prior_sampling``.  The perturbation-kernel
algebra the samplers need lives in :mod:`psld_b200.schedule` (host, float64); the training-only
methods (``perturb_data``, ``predict_x_from_eps``, ``likelihood_weighting``) are out of scope
(SURVEY.md §2 row 2).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .registry import register_module
from .schedule import PSLDSchedule, VPSchedule


@register_module(category="sde", name="psld_b200")
class PSLD(PSLDSchedule):
    def __init__(self, config):
        super().__init__(config)
        self.N = int(config.model.sde.n_timesteps)

    def __repr__(self):
        return (f"Initialized SDE with m_inv:{self.m_inv}, gamma: {self.gamma}, nu: {self.nu}, "
                f"Decomp mode: {self.decomp_mode}")

    @property
    def type(self):
        return f"psld-{self.mode}"

    def prior_sampling(self, shape):
        """CPU prior exactly like the reference (psld.py:366-370): cat[N(0,1), N(0,M)]."""
        p_x = torch.randn(*shape)
        p_m = torch.randn(*shape) * np.sqrt(self.m)
        return torch.cat([p_x, p_m], dim=1)

    def prior_sampling_device(self, shape, seed: int, device="cuda"):
        """Same law drawn on the GPU with Philox (removes the H2D bookend)."""
        B, Cc, H, W = shape
        u = torch.empty(B, 2 * Cc, H, W, dtype=torch.float32, device=device)
        L.check(L.lib().psld_prior_sample(L.ptr(u), float(np.sqrt(self.m)), int(seed), B,
                                          Cc * H * W, L.stream_ptr(u.device)), "psld_prior_sample")
        return u


@register_module(category="sde", name="vpsde_b200")
class VPSDE(VPSchedule):
    """The VP-SDE baseline (reference ``VPSDE``, vpsde.py:8-99): attribute surface used by sampling."""

    def __init__(self, config):
        super().__init__(config)
        self.N = int(config.model.sde.n_timesteps)

    @property
    def type(self):
        return "vpsde"

    def prior_sampling(self, shape):
        return torch.randn(*shape)
