"""String-keyed plug-in registry, the reference's drop-in mechanism.

The reference resolves the sampler / score network / SDE by name through
``util.get_module(category, name)`` (reference ``main/util.py:33-62``; call sites
``main/eval/sample.py:38,50,55``).  This module keeps the same two functions (same
argument meaning, same ``ValueError`` behaviour) and :func:`install` publishes the B200
classes into the reference's own ``util._MODULES`` so that ``main/eval/sample.py`` and the
``scripts_psld`` scripts run unchanged apart from ``name=`` overrides.
"""
from __future__ import annotations

_MODULES: dict = {}


def register_module(category=None, name=None):
    """Decorator; raises ``ValueError`` on a duplicate explicit name (util.py:46-50)."""

    def _register(cls):
        cat = category
        if cat is None:
            cat = cls.__name__ if name is None else name
        bucket = _MODULES.setdefault(cat, {})
        key = cls.__name__ if name is None else name
        if name in bucket:
            raise ValueError(f"Already registered module with name: {key} in category: {category}")
        bucket[key] = cls
        return cls

    return _register


def get_module(category, name):
    module = _MODULES.get(category, dict()).get(name, None)
    if module is None:
        raise ValueError(f"No module named `{name}` found in category: `{category}`")
    return module


def install(reference_util=None, override: bool = False):
    """Publishes this package's classes into the reference registry.

    ``reference_util``: the reference's imported ``util`` module (default: ``import util``).
    New names (``sscs_sde_b200``, ``em_sde_b200``, ``ip_em_sde_b200``, ``cc_em_sde_b200``, ``bb_ode_b200``,
    ``ncsnpp_b200``, ``cfg_ncsnpp_b200``, ``psld_b200``) are always
    added; with ``override=True`` the reference's own names (``sscs_sde``, ``em_sde``,
    ``ncsnpp``) are re-pointed too, by writing ``util._MODULES[category][name]`` directly
    (``register_module`` would raise on the duplicate, util.py:46-50).
    """
    from . import guidance, ncsnpp, ode, samplers, sde  # noqa: F401  (registers into _MODULES)
    if reference_util is None:
        import util as reference_util  # the reference's main/util.py must be on sys.path
    reg = reference_util._MODULES
    for cat, bucket in _MODULES.items():
        dst = reg.setdefault(cat, {})
        for name, cls in bucket.items():
            dst[name] = cls
            if override and name.endswith("_b200") and cat != "sde":
                dst[name[: -len("_b200")]] = cls
    return reg
