"""Motion-encoder entry as one tcgen05 kernel.

The reference's ``PCBlock4_Deep_nopool_res.forward`` (``core/update.py:30-36``) starts with

    x = F.gelu(x + self.ffn1(x))          # ffn1 = Conv2d(C, 1.5 C, 1) -> GELU -> Conv2d(1.5 C, C, 1)

and ``SKMotionEncoder6_Deep_nopool_res.convc1`` applies it to the 324-channel correlation feature the lookup has just
written (``core/update.py:320,330``).  ``pcblock_ffn1(x, ffn1)`` computes that line with ``sf_pcblock_ffn1``;
``patch_motion_encoder(model)`` binds it into the PCBlocks of a StreamFlow model whose widths the kernel supports
(C <= 384: ``convc1``, ``convc2``, ``convf2``, ``conv`` of the motion encoder) -- an optional caller-side patch like
``patch_upsample``; everything after the first line of the block runs the reference's own modules.

Result dtype follows the reference under autocast: fp32 input -> fp32 (fp32 + fp16 promotes), fp16 input -> fp16.
Inference only.
"""
from __future__ import annotations

import types
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import StreamCorrError
from .corr import _on_device, _stream_ptr

MAX_CHANNELS, MAX_HIDDEN = 384, 512
_PACKED: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _ceil(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _pack(ffn1):
    """fp16 zero-padded copies of the two 1x1 convolutions (cached per module; refreshed when a weight changes)."""
    c1, c2 = ffn1[0], ffn1[2]
    params = (c1.weight, c1.bias, c2.weight, c2.bias)
    key = tuple((p.data_ptr(), p._version if not p.is_inference() else -1, p.device) for p in params)
    hit = _PACKED.get(ffn1)
    if hit is not None and hit[0] == key:
        return hit[1]
    Hd, C = int(c1.weight.shape[0]), int(c1.weight.shape[1])
    if tuple(c1.weight.shape[2:]) != (1, 1) or tuple(c2.weight.shape) != (C, Hd, 1, 1) or c1.bias is None or c2.bias is None:
        raise StreamCorrError("pcblock_ffn1: ffn1 must be Conv2d(C, H, 1) -> GELU -> Conv2d(H, C, 1) with biases")
    if C > MAX_CHANNELS or Hd > MAX_HIDDEN or C < 16:
        raise StreamCorrError(f"pcblock_ffn1: specialised for 16 <= C <= {MAX_CHANNELS} and hidden <= {MAX_HIDDEN} "
                              f"(got C={C}, hidden={Hd}); no generic fallback")
    dev = c1.weight.device
    Kp, Hp, N2 = _ceil(C, 64), _ceil(Hd, 128), _ceil(C, 16)
    with torch.no_grad():
        w1p = torch.zeros((Hp, Kp), dtype=torch.float16, device=dev)
        w1p[:Hd, :C] = c1.weight.detach().reshape(Hd, C)
        b1p = torch.zeros((Hp,), dtype=torch.float32, device=dev)
        b1p[:Hd] = c1.bias.detach().float()
        w2p = torch.zeros((N2, Hp), dtype=torch.float16, device=dev)
        w2p[:C, :Hd] = c2.weight.detach().reshape(C, Hd)
        b2 = c2.bias.detach().float().contiguous()
    packed = (w1p, b1p, w2p, b2, C, Hd)
    _PACKED[ffn1] = (key, packed)
    return packed


def pcblock_ffn1(x: torch.Tensor, ffn1) -> torch.Tensor:
    """``F.gelu(x + ffn1(x))`` for ``x [P, C, h, w]`` (fp32 or fp16, CUDA) and the reference's ``ffn1`` Sequential."""
    if x.dim() != 4 or not x.is_cuda:
        raise StreamCorrError(f"pcblock_ffn1: x must be a CUDA tensor [P, C, h, w], got {tuple(x.shape)}")
    if x.dtype not in (torch.float32, torch.float16):
        raise StreamCorrError(f"pcblock_ffn1: x must be fp32 or fp16, got {x.dtype}")
    _lib.require_no_grad("pcblock_ffn1", x, *[p for p in ffn1.parameters()])
    w1p, b1p, w2p, b2, C, Hd = _pack(ffn1)
    P, Cx, h, w = x.shape
    if Cx != C or w1p.device != x.device:
        raise StreamCorrError(f"pcblock_ffn1: x has {Cx} channels on {x.device}, ffn1 expects {C} on {w1p.device}")
    xc = x.detach()
    if not xc.is_contiguous():
        xc = xc.contiguous()
    dev = xc.device
    with _on_device(dev):
        out = torch.empty_like(xc)
        rc = _lib.lib().sf_pcblock_ffn1(xc.data_ptr(), _lib.torch_dtype_code(xc.dtype), w1p.data_ptr(), b1p.data_ptr(),
                                        w2p.data_ptr(), b2.data_ptr(), out.data_ptr(), _lib.torch_dtype_code(out.dtype),
                                        P, C, Hd, h * w, _stream_ptr(dev))
    _lib.check(rc, "sf_pcblock_ffn1")
    return out


def _pcblock_forward(self, x):
    """``PCBlock4_Deep_nopool_res.forward`` (core/update.py:30-36) with its first line on ``sf_pcblock_ffn1``."""
    x = pcblock_ffn1(x, self.ffn1)
    for conv in self.conv_list:
        x = F.gelu(x + conv(x))
    x = F.gelu(x + self.pw(x))
    return self.ffn2(x)


def patch_motion_encoder(model) -> list:
    """Bind the fused first stage into every supported PCBlock of ``model.update_block.encoder`` (``convc1``, ``convc2``,
    ``convf2``, ``conv``); returns the names patched.  ``unpatch_motion_encoder`` restores the reference forward."""
    enc = model.update_block.encoder
    done = []
    for name in ("convc1", "convc2", "convf2", "conv"):
        blk = getattr(enc, name, None)
        if blk is None or not hasattr(blk, "ffn1") or not hasattr(blk, "conv_list"):
            continue
        c1 = blk.ffn1[0]
        if c1.weight.shape[1] > MAX_CHANNELS or c1.weight.shape[0] > MAX_HIDDEN or c1.weight.shape[1] < 16:
            continue
        blk.forward = types.MethodType(_pcblock_forward, blk)
        done.append(name)
    return done


def unpatch_motion_encoder(model) -> None:
    enc = model.update_block.encoder
    for name in ("convc1", "convc2", "convf2", "conv"):
        blk = getattr(enc, name, None)
        if blk is not None and "forward" in blk.__dict__:
            del blk.__dict__["forward"]
