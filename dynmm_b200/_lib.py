"""ctypes binding of ``libdynmm_b200.so`` (the C ABI declared in
``include/dynmm_b200.h``).

There is deliberately NO fallback: if the library is missing or a call fails
the error is raised.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynmm_b200.so")


class DynmmError(RuntimeError):
    pass


class TileFlags(Structure):
    """Mirror of ``dynmm_tile_flags`` (include/dynmm_b200.h)."""
    _fields_ = [("flags", c_void_p), ("box_n", c_int32), ("box_h", c_int32), ("box_w", c_int32),
                ("tiles_h", c_int32), ("tiles_w", c_int32), ("need", c_int32)]


class ConvParams(Structure):
    """Mirror of ``dynmm_conv_params`` (include/dynmm_b200.h)."""
    _fields_ = [
        ("in_", c_void_p), ("weight", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
        ("residual", c_void_p), ("res_map", c_void_p), ("out", c_void_p), ("gated", c_void_p), ("gate", c_void_p),
        ("gated_slot", c_void_p), ("in_map", c_void_p), ("count", c_void_p),
        ("n", c_int32), ("n_in", c_int32),
        ("h_in", c_int32), ("w_in", c_int32), ("c_in", c_int32), ("in_ld", c_int32),
        ("h_out", c_int32), ("w_out", c_int32), ("c_out", c_int32), ("out_ld", c_int32),
        ("res_ld", c_int32), ("gated_ld", c_int32),
        ("kh", c_int32), ("kw", c_int32), ("stride_h", c_int32), ("stride_w", c_int32),
        ("pad_h", c_int32), ("pad_w", c_int32),
        ("relu", c_int32), ("tile_n", c_int32), ("max_ctas", c_int32), ("flags", c_int32),
        ("trace", c_void_p),
        ("in_flags", TileFlags), ("res_flags", TileFlags), ("out_flags", TileFlags),
    ]


class ConvPairParams(Structure):
    """Mirror of ``dynmm_conv_pair_params`` (include/dynmm_b200.h)."""
    _fields_ = [
        ("in_", c_void_p), ("w1", c_void_p), ("shift1", c_void_p), ("w2", c_void_p), ("shift2", c_void_p),
        ("residual", c_void_p), ("out", c_void_p), ("count", c_void_p), ("in_map", c_void_p), ("res_map", c_void_p),
        ("n", c_int32), ("n_in", c_int32), ("h", c_int32), ("w", c_int32),
        ("in_ld", c_int32), ("out_ld", c_int32), ("res_ld", c_int32), ("relu2", c_int32),
    ]


class ChainLayer(Structure):
    """Mirror of ``dynmm_chain_layer`` (include/dynmm_b200.h)."""
    _fields_ = [("weight", c_void_p), ("shift", c_void_p), ("taps_h", c_int32), ("relu", c_int32),
                ("residual", c_int32), ("store", c_int32)]


class ChainJob(Structure):
    """Mirror of ``dynmm_chain_job`` (include/dynmm_b200.h)."""
    _fields_ = [("image", c_void_p), ("in_", c_void_p), ("out", c_void_p), ("out_last", c_void_p), ("count", c_void_p),
                ("n", c_int32), ("n_layers", c_int32), ("count_settled", c_int32), ("reserved", c_int32)]


class ChainParams(Structure):
    """Mirror of ``dynmm_chain_params`` (include/dynmm_b200.h)."""
    _fields_ = [("jobs", ChainJob * 2), ("n_jobs", c_int32), ("h", c_int32), ("w", c_int32), ("c", c_int32),
                ("flags", c_void_p), ("scratch", c_void_p), ("scratch_bytes", c_longlong), ("trace", c_void_p)]


class WgradParams(Structure):
    """Mirror of ``dynmm_wgrad_params`` (include/dynmm_b200.h)."""
    _fields_ = [
        ("x", c_void_p), ("dy", c_void_p), ("dw", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_longlong),
        ("n", c_int32), ("h_in", c_int32), ("w_in", c_int32), ("c_in", c_int32), ("x_ld", c_int32),
        ("h_out", c_int32), ("w_out", c_int32), ("c_out", c_int32), ("dy_ld", c_int32),
        ("kh", c_int32), ("kw", c_int32), ("stride_h", c_int32), ("stride_w", c_int32),
        ("pad_h", c_int32), ("pad_w", c_int32), ("accumulate", c_int32), ("max_ctas", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/dynmm_b200.h declares
SIGNATURES = {
    "dynmm_abi_version": (c_int, []),
    "dynmm_last_error": (c_char_p, []),
    "dynmm_device_ok": (c_int, []),
    "dynmm_diffsoftmax_fwd": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_diffsoftmax_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dynmm_gate_plan": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_global_gate_workspace": (c_longlong, [c_int, c_int, c_int]),
    "dynmm_global_gate_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 7
                                 + [c_void_p, c_void_p, c_void_p]),
    "dynmm_global_gate_decide": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 7
                                 + [c_void_p, c_float, c_int] + [c_void_p] * 8),
    "dynmm_stem_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 6 + [c_void_p] * 4
                       + [c_void_p] * 3 + [c_void_p]),
    "dynmm_stem_gap_tiles": (c_longlong, [c_int, c_int, c_int]),
    "dynmm_stem_s2d_workspace": (c_longlong, [c_int, c_int, c_int]),
    "dynmm_stem_s2d_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_stem_s2d_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p] + [c_void_p] * 4
                           + [c_void_p, c_longlong] + [c_void_p] * 4 + [c_int, c_void_p, c_void_p]),
    "dynmm_gap_workspace": (c_longlong, [c_int, c_int]),
    "dynmm_gap_partial": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dynmm_gap_partial_split": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dynmm_se_mlp": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_se_gated_fuse": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong,
                                    c_int, c_int, c_void_p, c_void_p]),
    "dynmm_se_gated_fuse_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong,
                                          c_int, c_int, c_void_p, c_void_p]),
    "dynmm_conv_igemm_fwd": (c_int, [POINTER(ConvParams), c_void_p]),
    "dynmm_conv_igemm_fwd2": (c_int, [POINTER(ConvParams), POINTER(ConvParams), c_void_p]),
    "dynmm_conv_tile_grid": (c_int, [POINTER(ConvParams), POINTER(TileFlags)]),
    "dynmm_conv_direct_fwd": (c_int, [POINTER(ConvParams), c_void_p]),
    "dynmm_conv_pair_fwd": (c_int, [POINTER(ConvPairParams), c_void_p]),
    "dynmm_conv_chain_image_bytes": (c_longlong, [c_int]),
    "dynmm_conv_chain_build": (c_int, [POINTER(ChainLayer), c_int, c_int, c_void_p]),
    "dynmm_conv_chain_plan": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_longlong)]),
    "dynmm_conv_chain_fwd": (c_int, [POINTER(ChainParams), c_void_p]),
    "dynmm_conv_program_bytes": (c_longlong, [c_int]),
    "dynmm_conv_program_build": (c_int, [POINTER(ConvParams), POINTER(c_int32), c_int, c_void_p, c_longlong,
                                         POINTER(c_int32)]),
    "dynmm_conv_program_launch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_conv_wgrad_workspace": (c_longlong, [POINTER(WgradParams)]),
    "dynmm_conv_wgrad": (c_int, [POINTER(WgradParams), c_void_p]),
    "dynmm_conv_wgrad_direct": (c_int, [POINTER(WgradParams), c_void_p]),
    "dynmm_pack_conv_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dynmm_fold_pack_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 5 + [c_float, c_void_p, c_void_p,
                                                                                           c_void_p]),
    "dynmm_fold_pack_conv_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 5 + [c_float, c_void_p, c_void_p,
                                                                                                 c_void_p]),
    "dynmm_split_from_f32": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    "dynmm_upsample2x_dw3x3_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p]),
    "dynmm_upsample2x_dw3x3_ex": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dynmm_bilinear_resize_into": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p]),
    "dynmm_adaptive_avgpool_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dynmm_nearest_resize_into_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                                c_void_p]),
    "dynmm_fold_bn": (c_int, [c_int] + [c_void_p] * 5 + [c_float, c_void_p, c_void_p, c_void_p]),
    "dynmm_permute3d_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dynmm_channel_sum_workspace": (c_longlong, [c_longlong, c_int]),
    "dynmm_channel_sum": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p, c_longlong, c_int,
                                  c_void_p]),
    "dynmm_gated_add_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "dynmm_gated_add_f32_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "dynmm_gated_add_f32_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_void_p]),
    "dynmm_softgate_mix_fwd": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_void_p]),
    "dynmm_softgate_mix_bwd": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_int, c_int, c_int, POINTER(c_void_p),
                                       c_void_p, c_void_p]),
    "dynmm_compact_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_argmax_confusion": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dynmm_miou": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dynmm_ce2d_workspace": (c_longlong, [c_int, c_int, c_int, c_int]),
    "dynmm_ce2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_longlong,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "dynmm_ce2d_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dynmm_nchw_f32_to_nhwc_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dynmm_nhwc_bf16_to_nchw_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dynmm_upsample2x_dw3x3": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "dynmm_adaptive_avgpool": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dynmm_nearest_resize_into": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                          c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and set every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DynmmError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dynmm_b200/csrc`.  There is no CPU or PyTorch fallback for the CUDA path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.dynmm_abi_version() != 1:
        raise DynmmError("libdynmm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dynmm_last_error().decode("utf-8", "replace")
        raise DynmmError(f"{what or 'dynmm call'} failed ({rc}): {msg}")


def ptr(t) -> int | None:
    """Raw device pointer of a tensor (None passes NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_device() -> None:
    """Fail loudly when the CUDA path cannot run (no silent fallbacks)."""
    if not torch.cuda.is_available():
        raise DynmmError("dynmm_b200 kernels need a CUDA device (sm_100a); none is visible")
    if load().dynmm_device_ok() != 1:
        raise DynmmError("dynmm_b200 kernels are built for sm_100a (B200) only")
