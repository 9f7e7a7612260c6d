"""Caller-side bookends of the sampling path: checkpoint loading and the
image writer's arithmetic, so that pretrained PSLD weights can be sampled end to end.

Reference: ``SDEWrapper.load_from_checkpoint`` (``main/eval/sample.py:62-69``; Lightning state
dict with prefixes ``score_fn.`` / ``ema_score_fn.``, chosen by ``evaluation.sample_from``,
``main/models/wrapper.py:30-31,44-46``) and ``SimpleImageWriter`` + ``save_as_images``
(``main/callbacks.py:88-124``, ``main/util.py:147-158``).
"""
from __future__ import annotations

import torch

from . import _lib as L


def select_score_fn_state(state_dict: dict, sample_from: str = "target") -> dict:
    """Picks the score network's tensors out of a Lightning ``SDEWrapper`` state dict:
    ``ema_score_fn.*`` when ``sample_from == "target"`` (wrapper.py:44-46), else ``score_fn.*``;
    returns them with the prefix stripped (keys ``all_modules.<i>...``)."""
    prefix = "ema_score_fn." if sample_from == "target" else "score_fn."
    out = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
    if not out:
        raise KeyError(f"no tensors with prefix `{prefix}` in the checkpoint")
    return out


def load_checkpoint(net, path: str, sample_from: str = "target", map_location="cpu", trust: bool = False):
    """Loads a reference ``.ckpt`` (Lightning: ``{"state_dict": ...}``) or a bare state dict into a
    :class:`psld_b200.NCSNpp`.  Parameter names/shapes are the reference's, so this is a plain
    ``load_state_dict`` (strict).

    Only tensors are needed, so the file is read with ``weights_only=True`` (no arbitrary pickle
    code runs).  A checkpoint that also pickles non-tensor objects (Lightning hyper-parameters,
    callbacks) needs ``trust=True`` to fall back to the unrestricted loader."""
    try:
        ckpt = torch.load(path, map_location=map_location, weights_only=True)
    except Exception as e:
        if not trust:
            raise RuntimeError(
                f"{path}: not loadable with weights_only=True ({type(e).__name__}: {str(e)[:200]}); "
                "pass trust=True to unpickle it without restrictions if you trust its origin") from e
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if any(k.startswith(("score_fn.", "ema_score_fn.")) for k in sd):
        sd = select_score_fn_state(sd, sample_from)
    net.load_state_dict(sd, strict=True)
    return net


def samples_to_uint8(state: torch.Tensor) -> torch.Tensor:
    """``[B, 2C, H, W]`` sampler state (fp64/fp32, on the GPU) -> ``[B, H, W, C]`` uint8 images,
    exactly the reference writer's arithmetic (drop momentum, ``x*0.5+0.5``, ``*255``, clip,
    truncate), in one kernel on the device instead of ``.cpu()`` + numpy."""
    if not state.is_cuda or state.dim() != 4 or state.shape[1] % 2:
        raise ValueError("expected a CUDA [B,2C,H,W] state tensor")
    B, C2, H, W = state.shape
    st = state.contiguous()
    out = torch.empty(B, H, W, C2 // 2, dtype=torch.uint8, device=state.device)
    L.check(L.lib().psld_quantize_images(L.ptr(st), L.dtype_code(st.dtype), L.ptr(out), B, C2 // 2,
                                         H * W, L.stream_ptr(state.device)), "psld_quantize_images")
    return out
