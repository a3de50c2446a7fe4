"""ctypes binding of libdowngan_b200.so (the C ABI in include/downgan_b200.h).

There is no fallback: if the shared library is missing or a call fails, the
error is raised to the caller.  The library is built in-tree by
``downgan_b200._build.build_library`` (``__graft_entry__.build()`` calls it).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdowngan_b200.so")

DG_FP32 = 0
DG_BF16 = 1


class DgError(RuntimeError):
    pass


class GeneratorConfig(C.Structure):
    _fields_ = [("filters", C.c_int), ("channels", C.c_int), ("n_predictands", C.c_int),
                ("num_res_blocks", C.c_int), ("num_upsample", C.c_int), ("coarse_dim", C.c_int),
                ("max_batch", C.c_int), ("precision", C.c_int)]


class CriticConfig(C.Structure):
    _fields_ = [("coarse_dim", C.c_int), ("fine_dim", C.c_int), ("nc", C.c_int),
                ("max_batch", C.c_int), ("precision", C.c_int)]


class Hyper(C.Structure):
    _fields_ = [("gp_lambda", C.c_float), ("gamma", C.c_float), ("content_lambda", C.c_float),
                ("freq_sep", C.c_int), ("filter_size", C.c_int)]


_P = C.c_void_p


class DpPeers(C.Structure):
    """dg_dp_peers (include/downgan_b200.h): peer-mapped gradient buckets and flag blocks of every rank."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("grad_ptrs", C.c_void_p * 8), ("flag_ptrs", C.c_void_p * 8),
                ("grad_multicast", C.c_void_p)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "dg_last_error": (C.c_char_p, []),
    "dg_abi_version": (C.c_int, []),
    "dg_has_tcgen05": (C.c_int, []),
    "dg_launch_count": (C.c_int64, []),
    "dg_profile": (C.c_int, [C.c_int]),
    "dg_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "dg_profile_report": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "dg_generator_create": (C.c_int, [C.POINTER(GeneratorConfig), C.POINTER(_P)]),
    "dg_generator_destroy": (C.c_int, [_P]),
    "dg_critic_create": (C.c_int, [C.POINTER(CriticConfig), C.POINTER(_P)]),
    "dg_critic_destroy": (C.c_int, [_P]),
    "dg_generator_param_count": (C.c_int64, [C.POINTER(GeneratorConfig)]),
    "dg_generator_param_offset": (C.c_int64, [C.POINTER(GeneratorConfig), C.c_int]),
    "dg_critic_param_count": (C.c_int64, [C.POINTER(CriticConfig)]),
    "dg_critic_param_offset": (C.c_int64, [C.POINTER(CriticConfig), C.c_int]),
    "dg_generator_pack": (C.c_int, [_P, _P, _P]),
    "dg_critic_pack": (C.c_int, [_P, _P, _P]),
    "dg_generator_fwd": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "dg_generator_bwd": (C.c_int, [_P, _P, _P, _P, _P]),
    "dg_critic_fwd": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "dg_critic_bwd": (C.c_int, [_P, _P, _P, _P, _P]),
    "dg_gp": (C.c_int, [_P, C.POINTER(Hyper), _P, _P, _P, C.c_int, _P, _P, _P, _P]),
    "dg_l1_loss": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P, _P, _P]),
    "dg_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_int, C.c_float, _P]),
    "dg_dp_allreduce_adam": (C.c_int, [C.POINTER(DpPeers), _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_int, C.c_float, C.c_uint, _P]),
    "dg_critic_pack_lazy": (C.c_int, [_P, _P]),
    "dg_critic_step": (C.c_int, [_P, _P, C.POINTER(Hyper), _P, _P, _P, C.c_int, _P, _P, _P]),
    "dg_generator_step": (C.c_int, [_P, _P, C.POINTER(Hyper), _P, _P, C.c_int, _P, _P, _P]),
    "dg_critic_defer_conv_grads": (C.c_int, [_P, C.c_int]),
    "dg_critic_step_finish": (C.c_int, [_P, _P, _P]),
    "dg_generator_lookahead": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "dg_generator_lookahead_first": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dg_generator_step_saved": (C.c_int, [_P, _P, C.POINTER(Hyper), _P, C.c_int, _P, _P, _P]),
    "dg_critic_step_fake": (C.c_int, [_P, _P, C.POINTER(Hyper), C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "dg_generator_activation": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "dg_critic_activation": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dg_generator_trunk_fwd": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "dg_generator_trunk_bwd": (C.c_int, [_P, _P, _P, _P, _P]),
    "dg_metrics": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "dg_gather_rows": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64, _P, _P]),
    "dg_lowpass": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dg_conv3x3_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_int, _P]),
    "dg_conv3x3_dgrad": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dg_conv3x3_wgrad": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}

EXPORTS = tuple(_SIGNATURES.keys())
PROFILE_CLASSES = ("conv_direct", "wgrad_direct", "conv_tcgen05", "wgrad_tcgen05", "dense_block_tcgen05", "linear",
                   "l1_loss", "gp_norm", "adam", "interpolate", "layout")

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises DgError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DgError(
            f"{LIB_PATH} is missing: build it with `python -m downgan_b200._build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # DG_TUNE="key=value,...": kernel-selection switches for A/B measurements (dg_set_tuning)
    for kv in filter(None, os.environ.get("DG_TUNE", "").split(",")):
        k, v = kv.split("=")
        lib.dg_set_tuning(int(k), int(v))
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().dg_last_error()
        raise DgError(f"libdowngan_b200 error {status}: {msg.decode() if msg else '?'}")


def ptr(t) -> Optional[int]:
    """data_ptr of a CUDA float32 contiguous tensor (or None)."""
    if t is None:
        return None
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise DgError("expected a contiguous float32 CUDA tensor, got "
                      f"{type(t).__name__} {getattr(t, 'dtype', None)} {getattr(t, 'device', None)}")
    return t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
