"""ctypes binding of libunpp.so (the C ABI declared in include/unpp.h).

There is deliberately no fallback: if the shared library is missing or an entry point fails, the
caller gets an exception — the product path never degrades to PyTorch/CPU kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UNPP_LIB") or os.path.join(HERE, "libunpp.so")  # UNPP_LIB: an experimental build of the same library

UNPP_MAX_SRC = 6
MODE_CONV, MODE_DECONV = 0, 1


class UnppError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("nsrc", C.c_int32),
        ("src", C.c_void_p * UNPP_MAX_SRC),
        ("src_C", C.c_int32 * UNPP_MAX_SRC),
        ("src_step", C.c_int32 * UNPP_MAX_SRC),
        ("src_oy", C.c_int32 * UNPP_MAX_SRC),
        ("src_ox", C.c_int32 * UNPP_MAX_SRC),
        ("taps", C.c_int32),
        ("n_total", C.c_int32),
        ("n_tile", C.c_int32),
        ("wpacked", C.c_void_p),
        ("bias", C.c_void_p),
        ("mode", C.c_int32),
        ("relu", C.c_int32),
        ("out", C.c_void_p),
        ("head_w", C.c_void_p),
        ("head_b", C.c_void_p),
        ("heat", C.c_void_p),
        ("logit", C.c_void_p),
        ("head_classes", C.c_int32),
        ("drop_mask", C.c_void_p),
        ("drop_scale", C.c_float),
        ("addend", C.c_void_p),
        ("relu_mask_src", C.c_void_p),
        ("stats_partial", C.c_void_p),
        ("stats_aux", C.c_void_p),
        ("aux_mean", C.c_void_p),
        ("aux_istd", C.c_void_p),
        ("block2x2", C.c_int32),
        ("lowres_src", C.c_void_p),
        ("lowres_wpacked", C.c_void_p),
        ("lowres_C", C.c_int32),
        ("bias_classes", C.c_int32),
        ("pooled", C.c_void_p),
    ]


class PackArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("scale", C.c_void_p),
        ("kind", C.c_int32), ("src_O", C.c_int32), ("src_I", C.c_int32), ("taps", C.c_int32),
        ("n_total", C.c_int32), ("n_tile", C.c_int32), ("n_begin", C.c_int32),
        ("k_begin", C.c_int32), ("k_count", C.c_int32), ("k8_total", C.c_int32), ("k_dst8", C.c_int32),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("nsrc", C.c_int32),
        ("src", C.c_void_p * UNPP_MAX_SRC),
        ("src_C", C.c_int32 * UNPP_MAX_SRC),
        ("dz", C.c_void_p),
        ("cout", C.c_int32), ("dz_step", C.c_int32), ("dz_oy", C.c_int32), ("dz_ox", C.c_int32),
        ("taps", C.c_int32),
        ("partial", C.c_void_p),
    ]


class ReduceJob(C.Structure):
    _fields_ = [
        ("partial", C.c_void_p), ("dst", C.c_void_p),
        ("stride", C.c_int64), ("s_co", C.c_int64), ("s_ci", C.c_int64), ("s_tap", C.c_int64),
        ("nparts", C.c_int32), ("taps", C.c_int32), ("cin_total", C.c_int32), ("cout", C.c_int32), ("ci_begin", C.c_int32), ("ci_count", C.c_int32),
        ("scale", C.c_float), ("block_end", C.c_int32),
    ]


class RefConvArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("nsrc", C.c_int32),
        ("src", C.c_void_p * UNPP_MAX_SRC),
        ("src_C", C.c_int32 * UNPP_MAX_SRC),
        ("weight", C.c_void_p), ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("cout", C.c_int32), ("taps", C.c_int32), ("relu", C.c_int32), ("sigmoid", C.c_int32),
        ("out", C.c_void_p),
    ]


class OptimArgs(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("final_lr", C.c_float), ("gamma", C.c_float), ("base_lr", C.c_float), ("grad_scale", C.c_float),
    ]


class OptimTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("state1", C.c_void_p), ("state2", C.c_void_p), ("n", C.c_int64), ("block_end", C.c_int32), ("reserved", C.c_int32)]


OPT_ADAMW, OPT_ADAM, OPT_ADABOUND, OPT_SGD, OPT_SGDW = range(5)


# Every symbol include/unpp.h declares, with its ctypes signature (tests check the export list).
_SIGNATURES = {
    "unpp_last_error": (C.c_char_p, []),
    "unpp_version": (C.c_int, []),
    "unpp_num_sms": (C.c_int, []),
    "unpp_conv_tc": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "unpp_conv_grid": (C.c_int, [C.POINTER(ConvArgs)]),
    "unpp_pack_weights": (C.c_int, [C.POINTER(PackArgs), C.c_void_p]),
    "unpp_pack_weights_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "unpp_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_bn_fold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unpp_compose_deconv_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unpp_u8_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_maxpool2x2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_argmax_peaks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unpp_argmax_peaks_split": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "unpp_argmax_splits": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "unpp_topk_peaks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unpp_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    "unpp_wgrad_grid": (C.c_int, [C.POINTER(WgradArgs)]),
    "unpp_wgrad_reduce": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long,
                                    C.c_float, C.c_void_p]),
    "unpp_reduce_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "unpp_sizeof_reduce_job": (C.c_int, []),
    "unpp_reduce_partials": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "unpp_bn_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unpp_bn_relu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_maxpool2x2_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_bn_bwd_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p]),
    "unpp_head_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_head_bwd_grid": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "unpp_adamw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                             C.c_int, C.c_float, C.c_void_p]),
    "unpp_adamw_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "unpp_dropout_mask": (C.c_int, [C.c_void_p, C.c_long, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p]),
    "unpp_create_heatmap": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "unpp_optim_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(OptimArgs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "unpp_bilinear_up2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_bilinear_up2x_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_ref_conv": (C.c_int, [C.POINTER(RefConvArgs), C.c_void_p]),
    "unpp_ref_deconv2x2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_ref_maxpool2x2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unpp_sizeof_ref_conv_args": (C.c_int, []),
    "unpp_optim_step_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(OptimArgs), C.c_int, C.c_void_p]),
    "unpp_sizeof_optim_tensor": (C.c_int, []),
    "unpp_sizeof_optim_args": (C.c_int, []),
    "unpp_sizeof_conv_args": (C.c_int, []),
    "unpp_sizeof_pack_args": (C.c_int, []),
    "unpp_sizeof_wgrad_args": (C.c_int, []),
}

_lib = None
_lock = threading.Lock()


def exported_symbols():
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """Load libunpp.so (once per process).  Raises UnppError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise UnppError(
                    f"{LIB_PATH} is missing: build it with `python -m unet_nested4tiny_objects_keypoints_b200._build` "
                    "(or __graft_entry__.build()). There is no non-CUDA fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError here = header/library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().unpp_last_error()
        raise UnppError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")
