"""CPU, build container only: the oracle against the UNMODIFIED reference imported from
/root/reference (skipped where the reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import psld_oracle as O
from oracle.ref_loader import NoiseBank, load_reference, reference_available, reference_time_grid
from oracle.weights import fill_state_dict, noise_bank, prior
from psld_b200 import NCSNpp, tiny_config

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


def test_state_dict_contract():
    """psld_b200.NCSNpp exposes exactly the reference's parameter names and shapes."""
    R = load_reference()
    for cfg in (tiny_config(),):
        ref = R.NCSNpp(cfg).state_dict()
        mine = NCSNpp(cfg).state_dict()
        assert list(ref.keys()) == list(mine.keys())
        for k in ref:
            assert tuple(ref[k].shape) == tuple(mine[k].shape), k


def test_forward_bit_exact():
    R = load_reference()
    cfg = tiny_config()
    net = R.NCSNpp(cfg).eval()
    sd = fill_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, 3)
    net.load_state_dict(sd)
    x = torch.randn(2, 6, 32, 32, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([0.9, 0.004])
    with torch.no_grad():
        y = net(x, t)
    assert torch.equal(y, O.ncsnpp_forward(cfg, sd, x, t))


@pytest.mark.parametrize("kind", ["sscs_sde", "em_sde"])
def test_sampler(kind):
    R = load_reference()
    cfg = tiny_config(sampler=kind, n_discrete_steps=12)
    cfg.data.image_size = 8
    fake = lambda u, t: torch.tanh(u) * t.view(-1, 1, 1, 1)
    sde = R.PSLD(cfg)
    ts, n = reference_time_grid(cfg)
    u0 = prior((2, 3, 8, 8), 0.5, 5)
    nb = noise_bank((2 if kind == "sscs_sde" else 1) * n, (2, 6, 8, 8), 6)
    with NoiseBank(nb):
        ref = R.get_module("samplers", kind)(cfg, sde, fake).sample(u0.clone(), ts, n)
    fn = O.sscs_sample if kind == "sscs_sde" else O.em_sample
    out = fn(cfg, fake, u0, O.time_grid(cfg)[0], n, nb)
    # 5e-7: the reference does its first step partly in fp32 (fp32 prior x python scalars), the
    # oracle promotes the prior to fp64 first (documented in oracle/psld_oracle.py)
    assert (out - ref).abs().max() <= 5e-7 * ref.abs().max()


def test_registry_install():
    """install() publishes the B200 classes into the reference registry (util.py:33-62)."""
    R = load_reference()
    import psld_b200
    reg = psld_b200.install(R.util)
    assert R.get_module("samplers", "sscs_sde_b200") is psld_b200.SSCSSampler
    assert R.get_module("score_fn", "ncsnpp_b200") is psld_b200.NCSNpp
    assert R.get_module("samplers", "sscs_sde") is not psld_b200.SSCSSampler   # no override
    keep = dict(reg["samplers"])
    try:
        psld_b200.install(R.util, override=True)
        assert R.get_module("samplers", "sscs_sde") is psld_b200.SSCSSampler
        assert R.get_module("score_fn", "ncsnpp") is psld_b200.NCSNpp
    finally:
        reg["samplers"].update(keep)
        reg["score_fn"]["ncsnpp"] = R.NCSNpp
