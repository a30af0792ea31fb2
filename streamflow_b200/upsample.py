"""Convex 8x flow upsampling as one kernel.

``upsample_flow(flow, mask, ratio=8)`` has the argument meaning and result of ``SKFlow_MF8.upsample_flow``
(``core/models/streamflow.py:82-93``): ``flow [N, 2, H, W]``, ``mask [N, 9*ratio*ratio, H, W]`` ->
``[N, 2, ratio*H, ratio*W]`` fp32.  The reference upsamples every refinement iteration although test mode returns
only the last one (``streamflow.py:139-144``); ``patch_upsample(model_cls)`` swaps the method in.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import StreamCorrError
from .corr import _on_device, _stream_ptr


def upsample_flow(flow, mask, ratio=8):
    if ratio != 8:
        raise StreamCorrError(f"upsample_flow is specialised for ratio 8 (got {ratio})")
    if flow.dim() != 4 or flow.shape[1] != 2 or not flow.is_cuda:
        raise StreamCorrError(f"flow must be a CUDA tensor [N, 2, H, W], got {tuple(flow.shape)}")
    N, _, H, W = flow.shape
    if tuple(mask.shape) != (N, 9 * ratio * ratio, H, W) or mask.device != flow.device:
        raise StreamCorrError(f"mask must be [{N}, {9 * ratio * ratio}, {H}, {W}] on the flow's device, "
                              f"got {tuple(mask.shape)}")
    _lib.require_no_grad("upsample_flow", flow, mask)
    f = flow.detach()
    if f.dtype != torch.float32 or not f.is_contiguous():
        f = f.float().contiguous()
    m = mask.detach()
    if not m.is_contiguous():
        m = m.contiguous()
    dev = f.device
    with _on_device(dev):
        out = torch.empty((N, 2, ratio * H, ratio * W), dtype=torch.float32, device=dev)
        rc = _lib.lib().sf_upsample_flow(f.data_ptr(), m.data_ptr(), _lib.torch_dtype_code(m.dtype), out.data_ptr(),
                                         N, H, W, ratio, _stream_ptr(dev))
    _lib.check(rc, "sf_upsample_flow")
    return out


def patch_upsample(model_cls) -> None:
    """Replace ``model_cls.upsample_flow`` (e.g. ``SKFlow_MF8``) by the fused kernel."""
    model_cls.upsample_flow = lambda self, flow, mask, ratio=8: upsample_flow(flow, mask, ratio)
