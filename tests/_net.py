"""Test helpers: seeded networks / samplers on the GPU and the oracle beside them."""
from __future__ import annotations

import numpy as np
import torch

from oracle import psld_oracle as O
from oracle.weights import fill_state_dict, noise_bank, prior
from psld_b200 import NCSNpp


def make_net(cfg, precision, seed=0, device="cuda"):
    net = NCSNpp(cfg).eval()
    net.set_precision(precision)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed)
    net.load_state_dict(sd)
    if device is not None:
        net = net.to(device)
    return net, sd


def sampler_inputs(cfg, B, n, kind, seed_p=1, seed_n=2):
    H = cfg.data.image_size
    sde = O.PSLDScalars(cfg)
    per = 2 if kind == "sscs_sde" else 1
    nb = noise_bank(per * n, (B, 6, H, H), seed_n)
    u0 = prior((B, 3, H, H), float(np.sqrt(sde.m)), seed_p)
    return u0, nb


def fake_score(u, t):
    """Same deterministic stand-in network as oracle/make_golden.py (device-agnostic)."""
    return (torch.tanh(u * 0.3) * 0.7 + 0.1 * torch.sin(torch.roll(u, 1, 1))) * t.view(-1, 1, 1, 1)


def vp_config(**ev):
    """VP-SDE baseline config of oracle/make_golden.py::vp_config (sample_uncond_vpsde.sh keys)."""
    from psld_b200 import tiny_config
    e = dict(sampler="em_sde", n_discrete_steps=40)
    e.update(ev)
    cfg = tiny_config(**e)
    cfg.model.sde.update(name="vpsde", beta_min=0.1, beta_max=20.0)
    cfg.model.score_fn.update(in_ch=3, out_ch=3)
    cfg.data.image_size = 8
    return cfg


def cc_config(**ev):
    """Classifier-guidance config of oracle/make_golden.py::cc_config (``config.clf.evaluation`` keys of
    the reference's class_cond_sample.py)."""
    from psld_b200 import tiny_config
    from psld_b200.config import Cfg
    e = dict(sampler="cc_em_sde", n_discrete_steps=40)
    e.update(ev)
    cfg = tiny_config(**e)
    cfg.data.image_size = 8
    cfg["clf"] = dict(evaluation=dict(label_to_sample=3, clf_temp=2.5))
    return Cfg(cfg)


def ode_config(tol=1e-5, **ev):
    """bb_ode config of oracle/make_golden.py::ode_config (sample_uncond_psld_ode.sh keys)."""
    from psld_b200 import tiny_config
    e = dict(sampler=dict(name="bb_ode", solver="RK45", rtol=tol, atol=tol))
    e.update(ev)
    cfg = tiny_config(**e)
    cfg.data.image_size = 8
    return cfg
