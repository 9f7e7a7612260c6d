"""ctypes binding of ``libpsld_b200.so`` (C ABI declared in ``include/psld_b200.h``).

The structures below mirror the header field for field.  There is NO fallback: if the
shared library cannot be loaded (or built with nvcc), every product entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# PSLD_B200_LIB: load another build of the library (kernel experiments next to the committed build)
LIB_PATH = os.environ.get("PSLD_B200_LIB") or os.path.join(HERE, "libpsld_b200.so")

# ---- constants (keep in sync with include/psld_b200.h) --------------------------------
VERSION = 100
OK, EINVAL, ECUDA, EUNSUPPORTED, ENUMERIC = 0, -1, -2, -3, -4
F32, BF16, F64, BF16S = 0, 1, 2, 3
NHWC, NCHW = 0, 1
STAGE_HALF_A, STAGE_SCORE, STAGE_HALF_B, STAGE_HALF_C = 1, 2, 4, 8
OP_LAYOUT, OP_TEMB, OP_GN, OP_FIR, OP_CONV, OP_ATTN, OP_ZERO, OP_AXPBY = 1, 2, 3, 4, 5, 6, 7, 8
ENGINE_SIMT, ENGINE_TC, ENGINE_TC_GN = 0, 1, 2
OP_NI, OP_NF, OP_NP = 28, 24, 10

# slot indices
(LAYOUT_N, LAYOUT_C, LAYOUT_HW, LAYOUT_DIR, LAYOUT_DTYPE, LAYOUT_CPAD, LAYOUT_CWRITE) = range(7)
(TEMB_NT, TEMB_NF, TEMB_EMB, TEMB_TOTALC, TEMB_LOGGED) = range(5)
(GN_N, GN_HW, GN_C1, GN_C2, GN_G, GN_SILU, GN_IN_DTYPE, GN_OUT_DTYPE, GN_NCHUNK,
 GN_AFFINE_ONLY) = range(10)
(FIR_N, FIR_H, FIR_W, FIR_C, FIR_UP, FIR_DOWN, FIR_PAD0, FIR_PAD1, FIR_KH, FIR_DTYPE,
 FIR_CACT) = range(11)
(CONV_N, CONV_H, CONV_W, CONV_C1, CONV_C2, CONV_COUT, CONV_KS, CONV_STRIDE, CONV_PAD, CONV_OH,
 CONV_OW, CONV_IN_LAYOUT, CONV_OUT_LAYOUT, CONV_IN_DTYPE, CONV_OUT_DTYPE, CONV_RES_DTYPE,
 CONV_TEMB_OFF, CONV_TEMB_BSTRIDE, CONV_GN_SILU, CONV_EXT_C1, CONV_EXT_C2) = range(21)
(ATTN_N, ATTN_HW, ATTN_C, ATTN_DTYPE, ATTN_PROJ) = range(5)


class HalfStep(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("a_xx", "a_xm", "a_mx", "a_mm", "c11", "c12", "c21", "c22")]


class ScoreStep(C.Structure):
    _fields_ = [("li11", C.c_float), ("li12", C.c_float), ("li21", C.c_float), ("li22", C.c_float),
                ("mode", C.c_int32), ("_pad", C.c_int32),
                ("k_x", C.c_double), ("k_m", C.c_double), ("m_inv", C.c_double),
                ("half_beta", C.c_double), ("gamma", C.c_double), ("nu", C.c_double),
                ("g2_x", C.c_double), ("g2_m", C.c_double), ("dt", C.c_double),
                ("gs_x", C.c_double), ("gs_m", C.c_double)]


class SscsCoeffs(C.Structure):
    _fields_ = [("half_a", HalfStep), ("half_b", HalfStep), ("half_c", HalfStep),
                ("score", ScoreStep)]


class InpaintStep(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("a_xx", "a_xm", "a_mx", "a_mm", "c11", "c12", "c21", "c22", "m0_std")] + \
               [("mean_only", C.c_int32), ("_pad", C.c_int32)]


class VpStep(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("half_beta", "g2", "neg_inv_std", "dt", "gs")]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("engine", C.c_int32),
                ("i", C.c_int32 * OP_NI), ("f", C.c_float * OP_NF),
                ("inp", C.c_void_p * OP_NP), ("out", C.c_void_p * OP_NP),
                ("aux", C.c_void_p)]


class SamplerDesc(C.Structure):
    _fields_ = [("sampler", C.c_int32), ("n_steps", C.c_int32), ("denoise", C.c_int32),
                ("state_dtype", C.c_int32), ("fuse_halves", C.c_int32), ("temb_op", C.c_int32),
                ("B", C.c_int64), ("chw", C.c_int64), ("seed", C.c_uint64),
                ("state", C.c_void_p), ("net_in", C.c_void_p), ("eps", C.c_void_p),
                ("time_table", C.c_void_p), ("noise", C.c_void_p),
                ("sscs", C.POINTER(SscsCoeffs)), ("em", C.POINTER(ScoreStep)),
                ("den", C.POINTER(ScoreStep)), ("record", C.c_void_p),
                ("sscs_dev", C.c_void_p), ("em_dev", C.c_void_p), ("step_counter", C.c_void_p)]


EXPORTS = {
    "psld_version": (C.c_int, []),
    "psld_last_error": (C.c_char_p, []),
    "psld_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "psld_sscs_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SscsCoeffs),
                                   C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64,
                                   C.c_void_p]),
    "psld_em_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.POINTER(ScoreStep), C.c_uint64, C.c_uint64,
                                 C.c_int64, C.c_int64, C.c_void_p]),
    "psld_em_update_guided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                        C.POINTER(ScoreStep), C.c_uint64, C.c_uint64, C.c_int64,
                                        C.c_int64, C.c_void_p]),
    "psld_reverse_drift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(ScoreStep),
                                     C.c_double, C.c_int64, C.c_int64, C.c_void_p]),
    "psld_vp_reverse_drift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VpStep),
                                        C.c_double, C.c_int64, C.c_void_p]),
    "psld_rk_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int64, C.c_void_p]),
    "psld_rk_error": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int,
                                C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_void_p]),
    "psld_inpaint_combine": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(InpaintStep), C.c_uint64,
                                       C.c_uint64, C.c_int64, C.c_int64, C.c_void_p]),
    "psld_vp_em_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.POINTER(VpStep), C.c_uint64, C.c_uint64, C.c_int64,
                                    C.c_void_p]),
    "psld_prior_sample": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64, C.c_int64, C.c_int64,
                                    C.c_void_p]),
    "psld_quantize_images": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_void_p]),
    "psld_axpby": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_int64,
                             C.c_void_p]),
    "psld_op_prepare": (C.c_int, [C.POINTER(Op)]),
    "psld_op_release": (C.c_int, [C.POINTER(Op)]),
    "psld_op_run": (C.c_int, [C.POINTER(Op), C.c_void_p]),
    "psld_program_run": (C.c_int, [C.POINTER(Op), C.c_int, C.c_void_p]),
    "psld_program_launches": (C.c_int, [C.POINTER(Op), C.c_int]),
    "psld_upfirdn2d": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int,
                                 C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "psld_sampler_run": (C.c_int, [C.POINTER(Op), C.c_int, C.POINTER(SamplerDesc), C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


class PsldError(RuntimeError):
    pass


def lib():
    """Loads (building with nvcc if it is absent) the CUDA library.  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        try:
            L = C.CDLL(LIB_PATH)
        except OSError as e:  # loud failure: there is no CPU / eager fallback
            raise PsldError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.psld_version() != VERSION:
            raise PsldError(f"libpsld_b200 version {L.psld_version()} != binding {VERSION}")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = lib().psld_last_error().decode(errors="replace")
        if rc == EUNSUPPORTED:
            raise NotImplementedError(f"{what}: {msg}")
        if rc == EINVAL:
            raise ValueError(f"{what}: {msg}")
        raise PsldError(f"{what}: {msg} (code {rc})")


def ptr(t):
    """Device (or host) data pointer of a torch tensor, or None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(dt):
    import torch
    return {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64}[dt]
