r"""
ctypes binding of ``libglue_b200.so`` (C ABI declared in ``include/glue_b200.h``).

There is no CPU fallback: if the shared library has not been built, or a kernel
is asked to run on a non-CUDA tensor, the caller gets a ``RuntimeError``.
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint32, c_void_p
from typing import Optional

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libglue_b200.so")

# Every symbol include/glue_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = (
    "glue_abi_version",
    "glue_arch",
    "glue_error_string",
    "glue_kernel_launches",
    "glue_graphconv_csr",
    "glue_csr_build_workspace_bytes",
    "glue_csr_build",
    "glue_graphenc_workspace_bytes",
    "glue_graphenc_head_fwd",
    "glue_graphenc_head_bwd",
    "glue_graphenc_head_bwd_impl",
    "glue_graphenc_head_fwd_impl",
    "glue_graphdec_workspace_bytes",
    "glue_graphdec_bce_fwd",
    "glue_graphdec_bwd",
    "glue_graphdec_bwd_scaled",
    "glue_graphdec_sort_incidence",
    "glue_graphdec_bwd_sorted",
    "glue_graph_nll_fwd",
    "glue_graph_nll_bwd",
    "glue_decoder_workspace_bytes",
    "glue_decoder_nll_fwd_bwd",
    "glue_decoder_nll_fwd",
    "glue_decoder_multi_workspace_bytes",
    "glue_decoder_nll_multi_fwd_bwd",
    "glue_decoder_nll_multi_fwd",
    "glue_decoder_log_prob",
    "glue_decoder_mean",
    "glue_rmsprop_flat",
    "glue_multi_copy",
    "glue_sample_graph_epoch",
    "glue_weighted_ce_fwd",
    "glue_kl_unit_normal_fwd",
    "glue_kl_unit_normal_bwd",
    "glue_paired_mean_cos_fwd",
    "glue_paired_mean_cos_bwd",
    "glue_act_bn_dropout_fwd",
    "glue_act_bn_dropout_bwd",
    "glue_mlp_hidden_fwd",
    "glue_mlp_heads_fwd",
    "glue_mlp_heads_bwd",
    "glue_mlp_hidden_bwd",
    "glue_tc_selftest",
    "glue_tc_set_trace",
)

FAMILY = {"NB": 0, "ZINB": 1, "Normal": 2, "ZILN": 3}


class DecoderArgs(ctypes.Structure):
    """Mirror of ``struct glue_decoder_args``."""

    _fields_ = [
        ("family", c_int32),
        ("B", c_int32),
        ("F", c_int32),
        ("K", c_int32),
        ("n_batches", c_int32),
        ("reduce", c_int32),
        ("impl", c_int32),
        ("grad_scale", c_float),
        ("u", c_void_p),
        ("v", c_void_p),
        ("vidx", c_void_p),
        ("b", c_void_p),
        ("row_order", c_void_p),
        ("l", c_void_p),
        ("x", c_void_p),
        ("scale_lin", c_void_p),
        ("bias", c_void_p),
        ("disp", c_void_p),
        ("zi_logits", c_void_p),
        ("wmat", c_void_p),
        ("row_weight", c_void_p),
        ("loss", c_void_p),
        ("grad_u", c_void_p),
        ("grad_v", c_void_p),
        ("grad_scale_lin", c_void_p),
        ("grad_bias", c_void_p),
        ("grad_disp", c_void_p),
        ("grad_zi_logits", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
    ]


MAX_TERMS = 3  # GLUE_DECODER_MAX_TERMS


class DecoderMultiArgs(ctypes.Structure):
    """Mirror of ``struct glue_decoder_multi_args``."""

    _fields_ = [
        ("shared", DecoderArgs),
        ("n_terms", c_int32),
        ("u", c_void_p * MAX_TERMS),
        ("row_weight", c_void_p * MAX_TERMS),
        ("grad_scale", c_float * MAX_TERMS),
        ("loss", c_void_p * MAX_TERMS),
        ("grad_u", c_void_p * MAX_TERMS),
    ]


ABI_VERSION = 6

_lib: Optional[ctypes.CDLL] = None


def _declare(lib: ctypes.CDLL) -> None:
    lib.glue_abi_version.restype = c_int32
    lib.glue_abi_version.argtypes = []
    lib.glue_arch.restype = c_char_p
    lib.glue_arch.argtypes = []
    lib.glue_error_string.restype = c_char_p
    lib.glue_error_string.argtypes = [c_int32]
    lib.glue_kernel_launches.restype = ctypes.c_ulonglong
    lib.glue_kernel_launches.argtypes = []

    lib.glue_graphconv_csr.restype = c_int32
    lib.glue_graphconv_csr.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]
    lib.glue_csr_build_workspace_bytes.restype = c_size_t
    lib.glue_csr_build_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.glue_csr_build.restype = c_int32
    lib.glue_csr_build.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
        c_void_p, c_size_t, c_void_p]

    lib.glue_graphenc_workspace_bytes.restype = c_size_t
    lib.glue_graphenc_workspace_bytes.argtypes = [c_int64, c_int32]
    lib.glue_graphenc_head_fwd.restype = c_int32
    lib.glue_graphenc_head_fwd.argtypes = [c_void_p] * 6 + [c_int64, c_int32] + [c_void_p] * 5 + [
        c_size_t, c_void_p]
    lib.glue_graphenc_head_bwd.restype = c_int32
    lib.glue_graphenc_head_bwd.argtypes = [c_void_p] * 8 + [c_int64, c_int32] + [c_void_p] * 6 + [
        c_size_t, c_void_p]

    lib.glue_graphenc_head_fwd_impl.restype = c_int32
    lib.glue_graphenc_head_fwd_impl.argtypes = [c_void_p] * 6 + [c_int64, c_int32] + [c_void_p] * 5 + [
        c_size_t, c_int32, c_void_p]
    lib.glue_graphenc_head_bwd_impl.restype = c_int32
    lib.glue_graphenc_head_bwd_impl.argtypes = [c_void_p] * 8 + [c_int64, c_int32] + [c_void_p] * 6 + [
        c_size_t, c_int32, c_void_p]

    lib.glue_graphdec_workspace_bytes.restype = c_size_t
    lib.glue_graphdec_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.glue_graphdec_bce_fwd.restype = c_int32
    lib.glue_graphdec_bce_fwd.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.glue_graphdec_bwd.restype = c_int32
    lib.glue_graphdec_bwd.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
        c_void_p, c_void_p, c_size_t, c_void_p]
    lib.glue_graphdec_bwd_scaled.restype = c_int32
    lib.glue_graphdec_bwd_scaled.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
        c_void_p, c_void_p, c_size_t, c_void_p]
    lib.glue_graphdec_sort_incidence.restype = c_int32
    lib.glue_graphdec_sort_incidence.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_void_p]
    lib.glue_graphdec_bwd_sorted.restype = c_int32
    lib.glue_graphdec_bwd_sorted.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
        c_void_p, c_void_p, c_size_t, c_void_p]
    lib.glue_graph_nll_fwd.restype = c_int32
    lib.glue_graph_nll_fwd.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
        c_void_p, c_void_p, c_int32, c_void_p, c_size_t, c_void_p]
    lib.glue_graph_nll_bwd.restype = c_int32
    lib.glue_graph_nll_bwd.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
        c_void_p, c_void_p, c_size_t, c_void_p]

    lib.glue_decoder_workspace_bytes.restype = c_size_t
    lib.glue_decoder_workspace_bytes.argtypes = [c_int32, c_int32, c_int32, c_int32, c_int32]
    for name in ("glue_decoder_nll_fwd_bwd", "glue_decoder_nll_fwd"):
        fn = getattr(lib, name)
        fn.restype = c_int32
        fn.argtypes = [POINTER(DecoderArgs), c_void_p]
    lib.glue_decoder_multi_workspace_bytes.restype = c_size_t
    lib.glue_decoder_multi_workspace_bytes.argtypes = [c_int32] * 6
    for name in ("glue_decoder_nll_multi_fwd_bwd", "glue_decoder_nll_multi_fwd"):
        fn = getattr(lib, name)
        fn.restype = c_int32
        fn.argtypes = [POINTER(DecoderMultiArgs), c_void_p]
    for name in ("glue_decoder_log_prob", "glue_decoder_mean"):
        fn = getattr(lib, name)
        fn.restype = c_int32
        fn.argtypes = [POINTER(DecoderArgs), c_void_p, c_void_p]


    lib.glue_tc_selftest.restype = c_int32
    lib.glue_tc_selftest.argtypes = [
        c_void_p, ctypes.c_uint32, c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
        ctypes.c_uint32, ctypes.c_uint32, POINTER(ctypes.c_uint32), POINTER(ctypes.c_uint32),
        ctypes.c_uint32, c_void_p, c_void_p]


    lib.glue_multi_copy.restype = c_int32
    lib.glue_multi_copy.argtypes = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.glue_weighted_ce_fwd.restype = c_int32
    lib.glue_weighted_ce_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_float, c_void_p,
                                         c_void_p, c_void_p]
    lib.glue_sample_graph_epoch.restype = c_int32
    lib.glue_sample_graph_epoch.argtypes = [
        c_uint32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32,
        c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.glue_rmsprop_flat.restype = c_int32
    lib.glue_rmsprop_flat.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                      c_float, c_void_p]
    lib.glue_kl_unit_normal_fwd.restype = c_int32
    lib.glue_kl_unit_normal_fwd.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p]
    lib.glue_kl_unit_normal_bwd.restype = c_int32
    lib.glue_kl_unit_normal_bwd.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p,
                                            c_void_p, c_void_p]
    lib.glue_paired_mean_cos_fwd.restype = c_int32
    lib.glue_paired_mean_cos_fwd.argtypes = [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p]
    lib.glue_paired_mean_cos_bwd.restype = c_int32
    lib.glue_paired_mean_cos_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32,
                                             c_void_p, c_void_p, c_void_p, c_void_p]
    lib.glue_act_bn_dropout_fwd.restype = c_int32
    lib.glue_act_bn_dropout_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_int64, c_int32, c_float, c_float, c_float, c_float, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p]
    lib.glue_act_bn_dropout_bwd.restype = c_int32
    lib.glue_act_bn_dropout_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                            c_int32, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.glue_mlp_hidden_fwd.restype = c_int32
    lib.glue_mlp_hidden_fwd.argtypes = [c_void_p, c_int64] + [c_void_p] * 8 + [c_int64, c_int32, c_int32] + [
        c_float] * 4 + [c_void_p] * 6
    lib.glue_mlp_heads_fwd.restype = c_int32
    lib.glue_mlp_heads_fwd.argtypes = [c_void_p, c_int64] + [c_void_p] * 4 + [c_int64, c_int32, c_int32] + [
        c_void_p] * 4
    lib.glue_mlp_heads_bwd.restype = c_int32
    lib.glue_mlp_heads_bwd.argtypes = [c_void_p, c_int64] + [c_void_p] * 3 + [c_int64, c_int32, c_int32] + [
        c_void_p] * 6
    lib.glue_mlp_hidden_bwd.restype = c_int32
    lib.glue_mlp_hidden_bwd.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_int32] + [c_void_p] * 6 + [
        c_int64, c_int32, c_int64, c_int32, c_float, c_float] + [c_void_p] * 6
    lib.glue_tc_set_trace.restype = None
    lib.glue_tc_set_trace.argtypes = [c_void_p]


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"glue_b200: native library {LIB_PATH} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C glue_b200/csrc`. There is no CPU fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        if handle.glue_abi_version() != ABI_VERSION:
            raise RuntimeError("glue_b200: ABI version mismatch, rebuild the native library")
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    """Turn a non-zero return code into ``RuntimeError`` (the ABI never throws)."""
    if code != 0:
        msg = lib().glue_error_string(code).decode()
        raise RuntimeError(f"glue_b200.{what} failed with code {code}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (``None`` -> NULL)."""
    return None if t is None else t.data_ptr()


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    """All given tensors must live on one CUDA device: the hot path has no CPU branch."""
    device = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "glue_b200 kernels run on CUDA (sm_100a) tensors only; got a "
                f"{t.device} tensor. There is no CPU fallback."
            )
        if device is None:
            device = t.device
        elif t.device != device:
            raise RuntimeError("glue_b200: tensors on different devices")
    if device is None:
        raise RuntimeError("glue_b200: no tensor given")
    return device


def current_stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
