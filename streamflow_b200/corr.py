"""Drop-in replacement for the reference's ``core/corr.py`` backed by libstreamcorr.so (sm_100a).

Same public surface as the reference (``core/corr.py:6-54``):

    CorrBlock(fmap1, fmap2, num_levels=4, radius=4)     # builds the 4-level correlation pyramid
    corr_fn(coords) -> Tensor[B, num_levels*(2r+1)**2, h, w]   fp32 contiguous
    CorrBlock.corr(fmap1, fmap2) -> Tensor[B, h, w, 1, h, w]
    attributes: num_levels, radius, corr_pyramid (list of [B*N, 1, h_l, w_l] tensors)

Differences that are invisible to the callers (``core/models/*.py``): the pyramid levels are row-padded
strided views (row pitch rounded up to 4 floats) and every call only enqueues kernels on the current CUDA
stream -- no host synchronisation (the reference does 4 CPU->GPU copies per lookup, core/corr.py:33).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import StreamCorrError

_DEFAULT_PRECISION = os.environ.get("STREAMCORR_PRECISION", "f16")


def _aligned_workspace(nbytes: int, device) -> tuple[torch.Tensor, int]:
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    ptr = (buf.data_ptr() + 1023) // 1024 * 1024
    return buf, ptr


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class CorrBlock:
    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, precision=None):
        if num_levels != _lib.NUM_LEVELS or radius != _lib.RADIUS:
            raise StreamCorrError(
                f"CorrBlock is specialised for num_levels={_lib.NUM_LEVELS}, radius={_lib.RADIUS} "
                f"(the only configuration the models use); got {num_levels}, {radius}")
        if fmap1.dim() != 4 or fmap1.shape != fmap2.shape:
            raise StreamCorrError(f"fmap1/fmap2 must be 4-D with equal shapes, got {tuple(fmap1.shape)} "
                                  f"and {tuple(fmap2.shape)}")
        if not fmap1.is_cuda or fmap1.device != fmap2.device:
            raise StreamCorrError("CorrBlock needs CUDA tensors on one device (no CPU fallback)")
        prec = precision if precision is not None else _DEFAULT_PRECISION
        if prec not in _lib.PRECISIONS:
            raise StreamCorrError(f"unknown precision {prec!r}; choose from {sorted(_lib.PRECISIONS)}")
        self.num_levels = num_levels
        self.radius = radius
        self.precision = prec
        # the model hands over fp32 maps (streamflow.py:107 `.float()`); other float dtypes are upcast
        f1 = fmap1 if fmap1.dtype == torch.float32 else fmap1.float()
        f2 = fmap2 if fmap2.dtype == torch.float32 else fmap2.float()
        B, D, h, w = f1.shape
        self._shape = (B, D, h, w)
        dev = f1.device
        L = _lib.lib()
        self._dims = [_lib.level_dims(h, w, l) for l in range(num_levels)]
        with torch.cuda.device(dev):
            self._levels = [torch.empty((B * h * w, hl * pitch), dtype=torch.float32, device=dev)
                            for (hl, wl, pitch) in self._dims]
            ws_bytes = L.sf_corr_workspace_bytes(B, D, h, w, _lib.PRECISIONS[prec])
            ws_buf, ws_ptr = _aligned_workspace(ws_bytes, dev)
            rc = L.sf_corr_build(f1.data_ptr(), f2.data_ptr(), B, D, h, w, _lib.i64_array(f1.stride()),
                                 _lib.i64_array(f2.stride()), _lib.ptr_array([t.data_ptr() for t in self._levels]),
                                 ws_ptr, ws_bytes, _lib.PRECISIONS[prec], _stream_ptr(dev))
        _lib.check(rc, "sf_corr_build")
        del ws_buf
        self.corr_pyramid = [
            t.as_strided((B * h * w, 1, hl, wl), (hl * pitch, hl * pitch, pitch, 1))
            for t, (hl, wl, pitch) in zip(self._levels, self._dims)
        ]
        self._level_ptrs = _lib.ptr_array([t.data_ptr() for t in self._levels])

    def __call__(self, coords):
        B, D, h, w = self._shape
        if coords.dim() != 4 or tuple(coords.shape) != (B, 2, h, w):
            raise StreamCorrError(f"coords must be [{B}, 2, {h}, {w}], got {tuple(coords.shape)}")
        dev = self._levels[0].device
        if coords.device != dev:
            raise StreamCorrError("coords live on a different device than the correlation pyramid")
        c = coords.detach()
        if c.dtype != torch.float32 or not c.is_contiguous():
            c = c.float().contiguous()
        side = 2 * self.radius + 1
        with torch.cuda.device(dev):
            out = torch.empty((B, self.num_levels * side * side, h, w), dtype=torch.float32, device=dev)
            rc = _lib.lib().sf_corr_lookup(self._level_ptrs, c.data_ptr(), out.data_ptr(), B, h, w, self.radius,
                                           self.num_levels, _stream_ptr(dev))
        _lib.check(rc, "sf_corr_lookup")
        return out

    @classmethod
    def from_dense_pyramid(cls, levels, radius=4):
        """Wrap an existing dense pyramid (list of [B*N, 1, h_l, w_l] fp32 CUDA tensors) -- used by the
        parity tests to exercise the lookup kernel on reference-built volumes."""
        self = cls.__new__(cls)
        self.num_levels, self.radius, self.precision = len(levels), radius, "external"
        if self.num_levels != _lib.NUM_LEVELS or radius != _lib.RADIUS:
            raise StreamCorrError("from_dense_pyramid: need 4 levels, radius 4")
        BN, _, h, w = levels[0].shape
        self._shape = (1, 0, h, w) if BN == h * w else (BN // (h * w), 0, h, w)
        self._dims = [_lib.level_dims(h, w, l) for l in range(self.num_levels)]
        self._levels = []
        for lv, (hl, wl, pitch) in zip(levels, self._dims):
            if tuple(lv.shape) != (BN, 1, hl, wl):
                raise StreamCorrError(f"level shape {tuple(lv.shape)} != {(BN, 1, hl, wl)}")
            buf = torch.zeros((BN, hl * pitch), dtype=torch.float32, device=lv.device)
            buf.view(BN, hl, pitch)[:, :, :wl] = lv[:, 0].float()
            self._levels.append(buf)
        self.corr_pyramid = [
            t.as_strided((BN, 1, hl, wl), (hl * pitch, hl * pitch, pitch, 1))
            for t, (hl, wl, pitch) in zip(self._levels, self._dims)
        ]
        self._level_ptrs = _lib.ptr_array([t.data_ptr() for t in self._levels])
        return self

    @staticmethod
    def corr(fmap1, fmap2):
        """All-pairs correlation volume [B, h, w, 1, h, w] (core/corr.py:46-54)."""
        blk = CorrBlock(fmap1, fmap2)
        B, _, h, w = blk._shape
        return blk.corr_pyramid[0].reshape(B, h, w, 1, h, w)


class CorrGroup:
    """The T-1 CorrBlocks of one clip looked up in ONE launch.

    ``CorrGroup(blocks)(coords_list)`` returns the ``[(B*(T-1)), 324, h, w]`` tensor that
    ``core/models/streamflow.py:132`` otherwise assembles with ``torch.stack`` + ``rearrange``; row order is
    ``b*(T-1) + t`` exactly as ``rearrange('B T C H W -> (B T) C H W')`` produces.
    """

    def __init__(self, blocks, out_dtype=torch.float32):
        blocks = list(blocks)
        if not 1 <= len(blocks) <= _lib.MAX_GROUPS:
            raise StreamCorrError(f"CorrGroup takes 1..{_lib.MAX_GROUPS} CorrBlocks")
        if any(b._shape != blocks[0]._shape for b in blocks):
            raise StreamCorrError("all CorrBlocks of a group must share one shape")
        if out_dtype not in (torch.float32, torch.float16):
            raise StreamCorrError("CorrGroup output dtype must be float32 or float16")
        self.blocks = blocks
        self.out_dtype = out_dtype
        self._level_ptrs = _lib.ptr_array([t.data_ptr() for b in blocks for t in b._levels])

    def __call__(self, coords_list):
        blocks = self.blocks
        G = len(blocks)
        B, _, h, w = blocks[0]._shape
        dev = blocks[0]._levels[0].device
        if len(coords_list) != G:
            raise StreamCorrError(f"expected {G} coordinate tensors")
        cs = []
        for c in coords_list:
            if tuple(c.shape) != (B, 2, h, w):
                raise StreamCorrError(f"coords must be [{B}, 2, {h}, {w}], got {tuple(c.shape)}")
            c = c.detach()
            cs.append(c if (c.dtype == torch.float32 and c.is_contiguous()) else c.float().contiguous())
        with torch.cuda.device(dev):
            out = torch.empty((B, G, 324, h, w), dtype=self.out_dtype, device=dev)
            if B == 1:
                outs = [out[0, g].data_ptr() for g in range(G)]
                rc = _lib.lib().sf_corr_lookup_group(
                    G, self._level_ptrs, _lib.ptr_array([c.data_ptr() for c in cs]), _lib.ptr_array(outs),
                    _lib.torch_dtype_code(self.out_dtype), B, h, w, 4, 4, _stream_ptr(dev))
                _lib.check(rc, "sf_corr_lookup_group")
            else:   # batch-strided destination: one launch per batch element keeps the (B T) row order
                for b in range(B):
                    lv = _lib.ptr_array([t[b * h * w:].data_ptr() for blk in blocks for t in blk._levels])
                    outs = [out[b, g].data_ptr() for g in range(G)]
                    rc = _lib.lib().sf_corr_lookup_group(
                        G, lv, _lib.ptr_array([c[b].data_ptr() for c in cs]), _lib.ptr_array(outs),
                        _lib.torch_dtype_code(self.out_dtype), 1, h, w, 4, 4, _stream_ptr(dev))
                    _lib.check(rc, "sf_corr_lookup_group")
        return out.view(B * G, 324, h, w)


def coords_grid(batch, ht, wd, device=None):
    """[B, 2, h, w] pixel grid, channel 0 = x, 1 = y (core/utils/utils.py:82-85)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack((xs, ys), dim=0).float()[None].expand(batch, -1, -1, -1)
