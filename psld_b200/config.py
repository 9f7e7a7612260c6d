"""Attribute-access config tree with the reference's Hydra key names.

The reference resolves every hot-path parameter through ``config.<group>.<key>``
attribute access on an OmegaConf ``DictConfig`` (reference
``main/configs/dataset/cifar10/cifar10_psld.yaml:1-100``; keys read by the hot
path: ``main/models/sde/psld.py:16-33``, ``main/models/score_fn/song_sde/ncsnpp.py:43-75``,
``main/models/wrapper.py:45-56``).  Everything in this package only ever uses
attribute access, so a ``DictConfig`` coming from ``main/eval/sample.py`` works
unchanged; :class:`Cfg` is the dependency-free stand-in used by ``bench.py`` and
the tests (omegaconf is not installed in the build image).
"""
from __future__ import annotations

import copy


class Cfg(dict):
    """dict with attribute access, recursively applied (DictConfig stand-in)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, Cfg):
                self[k] = Cfg(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - mirrors omegaconf behaviour
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})


_SCORE_FN_DEFAULTS = dict(
    name="ncsnpp", in_ch=6, out_ch=6, nonlinearity="swish", nf=128,
    ch_mult=[2, 2, 2], num_res_blocks=8, attn_resolutions=[16], dropout=0.15,
    resamp_with_conv=True, noise_cond=True, fir=True, fir_kernel=[1, 3, 3, 1],
    skip_rescale=True, resblock_type="biggan", progressive="none",
    progressive_input="residual", progressive_combine="sum",
    embedding_type="fourier", init_scale=0.0, fourier_scale=16,
)

_SDE_DEFAULTS = dict(
    name="psld", beta_min=8.0, beta_max=8.0, nu=4.01, gamma=0.01, kappa=0.04,
    decomp_mode="lower", numerical_eps=1e-9, n_timesteps=1000, is_augmented=True,
)

_EVAL_DEFAULTS = dict(
    sampler=dict(name="sscs_sde"), seed=0, n_discrete_steps=1000, denoise=True,
    eval_eps=1e-3, stride_type="uniform", use_pflow=False, sample_from="target",
    accelerator="gpu", devices=[0], n_samples=256, workers=1, batch_size=256,
    save_mode="image", sample_prefix="gpu", path_prefix="",
)


def make_config(image_size=32, num_channels=3, score_fn=None, sde=None, evaluation=None):
    """Build a ``config.dataset.diffusion``-shaped tree (reference ``sample.py:31``)."""
    sf = dict(_SCORE_FN_DEFAULTS); sf.update(score_fn or {})
    sd = dict(_SDE_DEFAULTS); sd.update(sde or {})
    ev = copy.deepcopy(_EVAL_DEFAULTS)
    for k, v in (evaluation or {}).items():
        if k == "sampler" and isinstance(v, str):
            ev["sampler"] = dict(name=v)
        else:
            ev[k] = v
    return Cfg(
        data=dict(name="synthetic", image_size=image_size, num_channels=num_channels, norm=True),
        model=dict(pl_module="sde_wrapper", score_fn=sf, sde=sd),
        training=dict(continuous=True, train_eps=1e-5),
        evaluation=ev,
    )


def tiny_config(**ev):
    """BASELINE.json configs[0]: tiny NCSN++ (nf=32, ch_mult=[1,2], 1 res block)."""
    e = dict(sampler="em_sde", n_discrete_steps=100, batch_size=8, n_samples=8)
    e.update(ev)
    return make_config(score_fn=dict(nf=32, ch_mult=[1, 2], num_res_blocks=1, init_scale=1.0),
                       evaluation=e)


def mid_config(**ev):
    """Mid-size net whose channel counts (64..256) are all tensor-core eligible."""
    e = dict(sampler="sscs_sde", n_discrete_steps=20, batch_size=4, n_samples=4)
    e.update(ev)
    return make_config(score_fn=dict(nf=64, ch_mult=[1, 2], num_res_blocks=2, init_scale=1.0),
                       evaluation=e)


def cifar10_config(**ev):
    """BASELINE.json configs[1]: CIFAR-10 SOTA NCSN++ (reference
    ``scripts_psld/sota/uncond/cifar10/sample_uncond_psld.sh:6-21``), SSCS 1000 steps."""
    e = dict(sampler="sscs_sde", n_discrete_steps=1000, batch_size=256, n_samples=256)
    e.update(ev)
    return make_config(evaluation=e)


def celeba64_config(**ev):
    """BASELINE.json configs[3]: CelebA-64 (reference
    ``scripts_psld/sota/uncond/celeba64/sample_uncond_psld.sh:6-21``)."""
    e = dict(sampler="sscs_sde", n_discrete_steps=1000, batch_size=64, n_samples=64)
    e.update(ev)
    return make_config(image_size=64,
                       score_fn=dict(ch_mult=[1, 2, 2, 2], num_res_blocks=4, dropout=0.1),
                       sde=dict(nu=4.005, gamma=0.005), evaluation=e)
