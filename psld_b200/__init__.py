"""psld_b200 — B200-native (sm_100a) reverse-time sampling hot path of PSLD.

Public surface mirrors the reference's plug-in interface for this path:

  registry ....... ``register_module`` / ``get_module`` / ``install``      (main/util.py:33-62)
  score_fn ....... ``NCSNpp(config)(u, t) -> eps``                         (song_sde/ncsnpp.py)
  sde ............ ``PSLD(config)``                                        (models/sde/psld.py)
  samplers ....... ``SSCSSampler`` / ``EulerMaruyamaSampler`` ``.sample``  (samplers/sde.py)

Everything computes through ``libpsld_b200.so`` (C ABI in ``include/psld_b200.h``); importing
this package does not need a GPU, using it does.
"""
from .config import Cfg, celeba64_config, cifar10_config, make_config, mid_config, tiny_config
from .registry import get_module, install, register_module
from .schedule import PSLDSchedule, StepTables, time_grid
from .sde import PSLD, VPSDE
from .ncsnpp import NCSNpp
from .guidance import ClassifierFreeGuidance
from .samplers import (ClassCondEulerMaruyamaSampler, EulerMaruyamaSampler, InpaintEulerMaruyamaSampler, Sampler,
                       SSCSSampler)
from .ode import BBODESampler
from .io import load_checkpoint, samples_to_uint8, select_score_fn_state

register_module(category="score_fn", name="ncsnpp_b200")(NCSNpp)

__all__ = [
    "Cfg", "make_config", "tiny_config", "mid_config", "cifar10_config", "celeba64_config",
    "register_module", "get_module", "install", "PSLDSchedule", "StepTables", "time_grid",
    "PSLD", "VPSDE", "NCSNpp", "ClassifierFreeGuidance", "SSCSSampler", "EulerMaruyamaSampler", "InpaintEulerMaruyamaSampler",
    "ClassCondEulerMaruyamaSampler", "BBODESampler", "Sampler",
    "load_checkpoint", "samples_to_uint8", "select_score_fn_state",
]
__version__ = "0.1.0"
