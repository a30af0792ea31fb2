"""ctypes binding of libstreamcorr.so (C ABI declared in include/streamcorr.h).

The library is the only compute backend: if it is missing or fails to load this module raises --
there is no CPU, eager-PyTorch or Triton fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstreamcorr.so")

# mirrors include/streamcorr.h
NUM_LEVELS = 4
RADIUS = 4
MAX_GROUPS = 8
PREC_F16, PREC_F16X2, PREC_FP32_SIMT, PREC_AUTO = 0, 1, 2, 3
DT_F32, DT_F16, DT_BF16 = 0, 1, 2
PRECISIONS = {"f16": PREC_F16, "f16x2": PREC_F16X2, "fp32": PREC_FP32_SIMT, "auto": PREC_AUTO}

KERNEL_LOOKUP, KERNEL_CORR_GEMM, KERNEL_GMA_AGGREGATE, KERNEL_GMA_STATS = 1, 2, 3, 4
KERNEL_CORR_PACK, KERNEL_GMA_PROJ, KERNEL_CORR_SIMT, KERNEL_UPSAMPLE = 5, 6, 8, 9

EXPORTS = [
    "sf_version", "sf_last_error", "sf_device_ok", "sf_launch_count", "sf_profile_kernel",
    "sf_debug_select_kernels", "sf_corr_level_dims", "sf_corr_workspace_bytes",
    "sf_corr_build", "sf_corr_lookup", "sf_corr_lookup_group", "sf_gma_npad", "sf_gma_e_elems",
    "sf_gma_workspace_bytes",
    "sf_gma_attention", "sf_gma_attention_qk", "sf_gma_aggregate", "sf_upsample_flow", "sf_pcblock_ffn1", "sf_debug_ffn1_trace",
]

_lib = None


class StreamCorrError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StreamCorrError(
            f"{LIB_PATH} not found: build it with `python -m streamflow_b200.build` "
            "(the package has no fallback path)")
    L = ctypes.CDLL(LIB_PATH)
    i64p = POINTER(c_int64)
    L.sf_version.restype = c_int
    L.sf_last_error.restype = c_char_p
    L.sf_device_ok.restype = c_int
    L.sf_launch_count.restype = c_int64
    L.sf_profile_kernel.argtypes = [c_int, c_void_p, c_void_p]
    L.sf_profile_kernel.restype = None
    L.sf_debug_select_kernels.argtypes = [c_int, c_int]
    L.sf_debug_select_kernels.restype = None
    L.sf_corr_level_dims.argtypes = [c_int64, c_int64, c_int, i64p, i64p, i64p, i64p]
    L.sf_corr_level_dims.restype = None
    L.sf_corr_workspace_bytes.argtypes = [c_int64, c_int64, c_int64, c_int64, c_int]
    L.sf_corr_workspace_bytes.restype = c_int64
    L.sf_corr_build.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, i64p, i64p,
                                POINTER(c_void_p), c_void_p, c_int64, c_int, c_void_p]
    L.sf_corr_build.restype = c_int
    L.sf_corr_lookup.argtypes = [POINTER(c_void_p), c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int,
                                 c_void_p]
    L.sf_corr_lookup.restype = c_int
    L.sf_corr_lookup_group.argtypes = [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_int,
                                       c_int64, c_int64, c_int64, c_int, c_int, c_void_p]
    L.sf_corr_lookup_group.restype = c_int
    L.sf_gma_npad.argtypes = [c_int64]
    L.sf_gma_npad.restype = c_int64
    L.sf_gma_e_elems.argtypes = [c_int64, c_int64]
    L.sf_gma_e_elems.restype = c_int64
    L.sf_gma_workspace_bytes.argtypes = [c_int64, c_int64, c_int64, c_int64]
    L.sf_gma_workspace_bytes.restype = c_int64
    L.sf_gma_attention.argtypes = [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_int,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    L.sf_gma_attention.restype = c_int
    L.sf_gma_attention_qk.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_float, c_int, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_void_p]
    L.sf_gma_attention_qk.restype = c_int
    L.sf_gma_aggregate.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64,
                                   c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]
    L.sf_gma_aggregate.restype = c_int
    L.sf_upsample_flow.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]
    L.sf_upsample_flow.restype = c_int
    L.sf_pcblock_ffn1.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                  c_int64, c_int64, c_int64, c_void_p]
    L.sf_pcblock_ffn1.restype = c_int
    L.sf_debug_ffn1_trace.argtypes = [c_void_p]
    L.sf_debug_ffn1_trace.restype = None
    _lib = L
    return L


def require_no_grad(what: str, *tensors) -> None:
    """The kernels are inference-only (no backward exists): refuse to run where autograd would expect a graph,
    instead of silently returning tensors without grad_fn (the reference modules are differentiable and trained
    through, train_mf.py:238-257)."""
    import torch
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise StreamCorrError(
            f"{what}: an input or parameter requires grad, but the B200 operators are inference-only (no backward "
            "kernels). Run under torch.no_grad() / torch.inference_mode(), or use the reference operators to train.")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().sf_last_error().decode("utf-8", "replace")
        raise StreamCorrError(f"{what} failed (code {rc}): {msg}")


def level_dims(h: int, w: int, level: int):
    """(h_l, w_l, tiles_y, tiles_x) of pyramid level `level`; a query image holds tiles_y*tiles_x*16 floats."""
    a, b, c, d = c_int64(), c_int64(), c_int64(), c_int64()
    lib().sf_corr_level_dims(h, w, level, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
    return a.value, b.value, c.value, d.value


def i64_array(values):
    return (c_int64 * len(values))(*[int(v) for v in values])


def ptr_array(values):
    return (c_void_p * len(values))(*[int(v) for v in values])


def torch_dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return DT_F32
    if dtype == torch.float16:
        return DT_F16
    if dtype == torch.bfloat16:
        return DT_BF16
    raise StreamCorrError(f"unsupported dtype {dtype}")
