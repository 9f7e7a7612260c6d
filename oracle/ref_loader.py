"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/``
and by the ``not gpu`` tests that pin the oracle restatement against the real
reference when it is present.  Nothing under ``psld_b200/`` imports this.

The reference needs three stubs to import on CPU without its absent dependencies
(SURVEY.md §8c / Appendix C):
  1. ``torch.utils.cpp_extension.load`` → no-op (``op/upfirdn2d.py:10``,
     ``op/fused_act.py:11`` JIT-compile at import; CPU tensors then take
     ``upfirdn2d_native``, ``op/upfirdn2d.py:146-149``);
  2. a stand-in ``torchdiffeq`` module (``samplers/ode.py:2``) whose ``odeint`` restates the
     ``scipy_solver`` wrapper of torchdiffeq 0.2.3 (the only method the reference calls);
  3. a bare ``models`` package so ``models/__init__.py`` (pytorch_lightning) is skipped.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("PSLD_REFERENCE_ROOT", "/root/reference")
_loaded = None


def scipy_solver_odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, **unused):
    """Restatement of the ONE code path of ``torchdiffeq==0.2.3`` (``Pipfile:8``; absent here and not
    part of the reference tree) that the reference uses: ``odeint(..., method="scipy_solver",
    options={"solver": ...})`` = ``ScipyWrapperODESolver`` (torchdiffeq/_impl/scipy_wrapper.py): the state
    is flattened to a float64-castable numpy vector, ``scipy.integrate.solve_ivp`` drives the solver
    with ``t_eval = t``, and every function evaluation converts ``(t, y)`` to tensors of ``y0``'s
    device AND dtype before calling ``func``; the solution comes back in ``y0``'s dtype."""
    import numpy as np
    import torch
    from scipy.integrate import solve_ivp
    assert method == "scipy_solver", method
    solver = (options or {}).get("solver", "LSODA")
    dtype, device, shape = y0.dtype, y0.device, y0.shape
    y0n = y0.detach().cpu().numpy().reshape(-1)

    def np_func(tt, y):
        tt = torch.tensor(tt).to(device, dtype)
        y = torch.reshape(torch.tensor(y).to(device, dtype), shape)
        with torch.no_grad():
            f = func(tt, y)
        return f.detach().cpu().numpy().reshape(-1)

    if t.numel() == 1:
        return torch.tensor(y0n)[None].to(device, dtype)
    tn = t.detach().cpu().numpy()
    sol = solve_ivp(np_func, t_span=[tn.min(), tn.max()], y0=y0n, t_eval=tn, method=solver, rtol=rtol,
                    atol=atol, max_step=float("inf"))
    out = torch.tensor(sol.y).T.to(device, dtype)
    return out.reshape(-1, *shape)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "main", "samplers"))


def load_reference():
    """Returns a namespace with PSLD, NCSNpp, get_module, samplers module, upfirdn2d_native."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    import torch  # noqa: F401
    import torch.utils.cpp_extension as ce

    main = os.path.join(REF_ROOT, "main")
    if main not in sys.path:
        sys.path.insert(0, main)
    real_load = ce.load
    ce.load = lambda *a, **k: types.SimpleNamespace()
    try:
        if "torchdiffeq" not in sys.modules:
            td = types.ModuleType("torchdiffeq")
            td.odeint = scipy_solver_odeint
            sys.modules["torchdiffeq"] = td
        if "models" not in sys.modules:
            pkg = types.ModuleType("models")
            pkg.__path__ = [os.path.join(main, "models")]
            sys.modules["models"] = pkg
        import samplers  # registers em_sde, sscs_sde, ...
        import samplers.sde as samplers_sde
        from models.sde.psld import PSLD
        from models.sde.vpsde import VPSDE  # noqa: F401  (registers sde/vpsde)
        from models.score_fn.song_sde.ncsnpp import NCSNpp
        from models.score_fn.song_sde import layerspp, up_or_down_sampling
        from models.score_fn.song_sde.op.upfirdn2d import upfirdn2d_native
        import util
    finally:
        ce.load = real_load
    _loaded = types.SimpleNamespace(
        PSLD=PSLD, VPSDE=VPSDE, NCSNpp=NCSNpp, samplers=samplers, samplers_sde=samplers_sde,
        get_module=util.get_module, util=util, layerspp=layerspp,
        up_or_down_sampling=up_or_down_sampling, upfirdn2d_native=upfirdn2d_native,
    )
    return _loaded


class NoiseBank:
    """Context manager: makes ``torch.randn_like`` (as seen by the reference samplers,
    which call it through the ``torch`` module global, ``samplers/sde.py:24,303,346``)
    pop pre-drawn tensors in call order.  Draws are cast to the dtype of the argument,
    as a real ``randn_like`` would return."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.i = 0

    def __enter__(self):
        import torch
        self._torch = torch
        self._orig = torch.randn_like

        def fake(x, *a, **k):
            if self.i >= len(self.draws):
                # the SSCS denoise step draws one tensor it discards (sde.py:346)
                self.i += 1
                return torch.zeros_like(x)
            z = self.draws[self.i]
            self.i += 1
            assert z.shape == x.shape, (z.shape, x.shape)
            return z.to(dtype=x.dtype, device=x.device)

        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        self._torch.randn_like = self._orig
        return False


def reference_time_grid(config, sde_T=1.0):
    """Time grid exactly as ``main/models/wrapper.py:51-54,101-114``."""
    import torch
    ev = config.evaluation
    n = ev.n_discrete_steps - 1 if ev.denoise else ev.n_discrete_steps
    t_final = sde_T - ev.eval_eps
    ts = torch.linspace(0, t_final, n + 1, dtype=torch.float64)
    if ev.stride_type == "quadratic":
        ts = t_final * torch.flip(1 - (ts / t_final) ** 2.0, dims=[0])
    return ts, n
