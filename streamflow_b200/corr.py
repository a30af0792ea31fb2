"""Drop-in replacement for the reference's ``core/corr.py`` backed by libstreamcorr.so (sm_100a).

Same public surface as the reference (``core/corr.py:6-54``):

    CorrBlock(fmap1, fmap2, num_levels=4, radius=4)     # builds the 4-level correlation pyramid
                                                        # (+ precision="auto" | "f16" | "f16x2" | "fp32")
    corr_fn(coords) -> Tensor[B, num_levels*(2r+1)**2, h, w]   fp32 contiguous
    CorrBlock.corr(fmap1, fmap2) -> Tensor[B, h, w, 1, h, w]
    attributes: num_levels, radius, corr_pyramid (list of [B*N, 1, h_l, w_l] tensors)

Differences that are invisible to the callers (``core/models/*.py``): the pyramid levels live in a 4x4-tiled
layout (``include/streamcorr.h``) -- ``corr_pyramid`` materialises reference-shaped copies on first access,
nothing on the hot path reads it -- and every call only enqueues kernels on the current CUDA stream, with no host
synchronisation (the reference does 4 CPU->GPU copies per lookup, core/corr.py:33).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import StreamCorrError

# "auto" (include/streamcorr.h SF_PREC_AUTO): the device-side absmax pass decides -- fp16-representable feature maps (the
# model's mixed-precision path) take the single-product fast path, anything else the fp32-faithful three-product path
_DEFAULT_PRECISION = os.environ.get("STREAMCORR_PRECISION", "auto")


def _aligned_workspace(nbytes: int, device) -> tuple[torch.Tensor, int]:
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    ptr = (buf.data_ptr() + 1023) // 1024 * 1024
    return buf, ptr


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` costs ~10 us per call; skip it when `dev` is already current."""

    __slots__ = ("_ctx",)

    def __init__(self, dev):
        self._ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


_DIMS_CACHE: dict = {}


def _level_dims(h: int, w: int):
    key = (h, w)
    if key not in _DIMS_CACHE:
        _DIMS_CACHE[key] = [_lib.level_dims(h, w, l) for l in range(_lib.NUM_LEVELS)]
    return _DIMS_CACHE[key]


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
        _lib.require_no_grad("CorrBlock", fmap1, fmap2)
        prec = precision if precision is not None else _DEFAULT_PRECISION
        if prec not in _lib.PRECISIONS:
            raise StreamCorrError(f"unknown precision {prec!r}; choose from {sorted(_lib.PRECISIONS)}")
        self.num_levels = num_levels
        self.radius = radius
        self.precision = prec
        # the model hands over fp32 maps (streamflow.py:107 `.float()`); other float dtypes are upcast
        f1 = fmap1.detach() if fmap1.dtype == torch.float32 else fmap1.detach().float()
        f2 = fmap2.detach() if fmap2.dtype == torch.float32 else fmap2.detach().float()
        B, D, h, w = f1.shape
        self._shape = (B, D, h, w)
        dev = f1.device
        L = _lib.lib()
        self._dims = _level_dims(h, w)
        with _on_device(dev):
            self._levels = [torch.empty((B * h * w, th * tw * 16), dtype=torch.float32, device=dev)
                            for (hl, wl, th, tw) in self._dims]
            ws_bytes = L.sf_corr_workspace_bytes(B, D, h, w, _lib.PRECISIONS[prec])
            ws_buf, ws_ptr = _aligned_workspace(ws_bytes, dev)
            rc = L.sf_corr_build(f1.data_ptr(), f2.data_ptr(), B, D, h, w, _lib.i64_array(f1.stride()),
                                 _lib.i64_array(f2.stride()), _lib.ptr_array([t.data_ptr() for t in self._levels]),
                                 ws_ptr, ws_bytes, _lib.PRECISIONS[prec], _stream_ptr(dev))
        _lib.check(rc, "sf_corr_build")
        del ws_buf
        self._pyramid_cache = None
        self._level_ptrs = _lib.ptr_array([t.data_ptr() for t in self._levels])

    @property
    def corr_pyramid(self):
        """Reference-shaped pyramid: list of [B*N, 1, h_l, w_l] fp32 tensors (un-tiled copies, built lazily)."""
        if self._pyramid_cache is None:
            out = []
            for t, (hl, wl, th, tw) in zip(self._levels, self._dims):
                img = t.view(-1, th, tw, 4, 4).permute(0, 1, 3, 2, 4).reshape(-1, th * 4, tw * 4)
                out.append(img[:, None, :hl, :wl].contiguous())
            self._pyramid_cache = out
        return self._pyramid_cache

    def __call__(self, coords):
        B, D, h, w = self._shape
        if coords.dim() != 4 or tuple(coords.shape) != (B, 2, h, w):
            raise StreamCorrError(f"coords must be [{B}, 2, {h}, {w}], got {tuple(coords.shape)}")
        dev = self._levels[0].device
        if coords.device != dev:
            raise StreamCorrError("coords live on a different device than the correlation pyramid")
        _lib.require_no_grad("CorrBlock.__call__", coords)
        c = coords.detach()
        if c.dtype != torch.float32 or not c.is_contiguous():
            c = c.float().contiguous()
        side = 2 * self.radius + 1
        with _on_device(dev):
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
        self._dims = _level_dims(h, w)
        self._levels = []
        for lv, (hl, wl, th, tw) in zip(levels, self._dims):
            if tuple(lv.shape) != (BN, 1, hl, wl):
                raise StreamCorrError(f"level shape {tuple(lv.shape)} != {(BN, 1, hl, wl)}")
            pad = torch.zeros((BN, th * 4, tw * 4), dtype=torch.float32, device=lv.device)
            pad[:, :hl, :wl] = lv[:, 0].float()
            buf = pad.view(BN, th, 4, tw, 4).permute(0, 1, 3, 2, 4).reshape(BN, th * tw * 16).contiguous()
            self._levels.append(buf)
        self._pyramid_cache = None
        self._level_ptrs = _lib.ptr_array([t.data_ptr() for t in self._levels])
        return self

    @classmethod
    def _from_tiled_levels(cls, levels, shape, precision, radius=4):
        """Wrap level buffers (tiled layout, [B*N, th*tw*16] fp32 views) that a batched build already filled."""
        self = cls.__new__(cls)
        self.num_levels, self.radius, self.precision = _lib.NUM_LEVELS, radius, precision
        self._shape = tuple(shape)
        self._dims = _level_dims(shape[2], shape[3])
        self._levels = list(levels)
        self._pyramid_cache = None
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

    @classmethod
    def from_fmaps(cls, fmaps, radius=4, precision=None, out_dtype=torch.float32):
        """Build the T-1 pyramids of a clip from its frame features ``fmaps [B, T, D, h, w]`` (any strides) and
        return their group -- what ``core/models/streamflow.py:110`` does with T-1 ``CorrBlock(fmaps[:, i],
        fmaps[:, i+1])`` calls.  For one clip (B = 1) the pairs are the batch of ONE build: frames 0..T-2 are packed
        as queries, frames 1..T-1 as pooled targets, and a single persistent GEMM launch writes all pyramids.  For
        B > 1 every pair index is one build batched over the clips.  ``group.blocks[i]`` are ordinary CorrBlocks."""
        if radius != _lib.RADIUS:
            raise StreamCorrError(f"CorrGroup is specialised for radius={_lib.RADIUS}; got {radius}")
        if fmaps.dim() != 5 or fmaps.shape[1] < 2:
            raise StreamCorrError(f"fmaps must be [B, T >= 2, D, h, w], got {tuple(fmaps.shape)}")
        if not fmaps.is_cuda:
            raise StreamCorrError("CorrGroup needs CUDA tensors (no CPU fallback)")
        B, T, D, h, w = fmaps.shape
        G = T - 1
        if G > _lib.MAX_GROUPS:
            raise StreamCorrError(f"CorrGroup takes 1..{_lib.MAX_GROUPS} pairs, got {G}")
        prec = precision if precision is not None else _DEFAULT_PRECISION
        if prec not in _lib.PRECISIONS:
            raise StreamCorrError(f"unknown precision {prec!r}; choose from {sorted(_lib.PRECISIONS)}")
        _lib.require_no_grad("CorrGroup.from_fmaps", fmaps)
        fm = fmaps.detach()
        if fm.dtype != torch.float32:
            fm = fm.float()
        dev = fm.device
        L = _lib.lib()
        dims = _level_dims(h, w)
        N = h * w
        code = _lib.PRECISIONS[prec]
        with _on_device(dev):
            # block-major storage: pair g owns rows [g * B * N, (g + 1) * B * N) of every level
            store = [torch.empty((G, B * N, th * tw * 16), dtype=torch.float32, device=dev)
                     for (hl, wl, th, tw) in dims]
            stream = _stream_ptr(dev)
            if B == 1:
                f1, f2 = fm[0, :-1], fm[0, 1:]
                ws_bytes = L.sf_corr_workspace_bytes(G, D, h, w, code)
                ws_buf, ws_ptr = _aligned_workspace(ws_bytes, dev)
                rc = L.sf_corr_build(f1.data_ptr(), f2.data_ptr(), G, D, h, w, _lib.i64_array(f1.stride()),
                                     _lib.i64_array(f2.stride()), _lib.ptr_array([t.data_ptr() for t in store]),
                                     ws_ptr, ws_bytes, code, stream)
                _lib.check(rc, "sf_corr_build")
            else:
                ws_bytes = L.sf_corr_workspace_bytes(B, D, h, w, code)
                ws_buf, ws_ptr = _aligned_workspace(ws_bytes, dev)
                for g in range(G):
                    f1, f2 = fm[:, g], fm[:, g + 1]
                    rc = L.sf_corr_build(f1.data_ptr(), f2.data_ptr(), B, D, h, w, _lib.i64_array(f1.stride()),
                                         _lib.i64_array(f2.stride()),
                                         _lib.ptr_array([t[g].data_ptr() for t in store]), ws_ptr, ws_bytes, code,
                                         stream)
                    _lib.check(rc, "sf_corr_build")
            del ws_buf
        blocks = [CorrBlock._from_tiled_levels([t[g] for t in store], (B, D, h, w), prec, radius) for g in range(G)]
        return cls(blocks, out_dtype=out_dtype)

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
            _lib.require_no_grad("CorrGroup.__call__", c)
            c = c.detach()
            cs.append(c if (c.dtype == torch.float32 and c.is_contiguous()) else c.float().contiguous())
        L = _lib.lib()
        code = _lib.torch_dtype_code(self.out_dtype)
        with _on_device(dev):
            out = torch.empty((B, G, 324, h, w), dtype=self.out_dtype, device=dev)
            base, per = out.data_ptr(), 324 * h * w * out.element_size()
            stream = _stream_ptr(dev)
            if B == 1:
                rc = L.sf_corr_lookup_group(G, self._level_ptrs, _lib.ptr_array([c.data_ptr() for c in cs]),
                                            _lib.ptr_array([base + g * per for g in range(G)]), code, 1, h, w, 4,
                                            4, stream)
                _lib.check(rc, "sf_corr_lookup_group")
            else:   # batch-strided destination: one launch per batch element keeps the (B T) row order
                img = [th * tw * 16 * 4 for (hl, wl, th, tw) in blocks[0]._dims]
                for b in range(B):
                    lv = _lib.ptr_array([t.data_ptr() + b * h * w * img[l]
                                         for blk in blocks for l, t in enumerate(blk._levels)])
                    rc = L.sf_corr_lookup_group(G, lv, _lib.ptr_array([c.data_ptr() + b * 2 * h * w * 4 for c in cs]),
                                                _lib.ptr_array([base + (b * G + g) * per for g in range(G)]), code,
                                                1, h, w, 4, 4, stream)
                    _lib.check(rc, "sf_corr_lookup_group")
        return out.view(B * G, 324, h, w)


def coords_grid(batch, ht, wd, device=None):
    """[B, 2, h, w] pixel grid, channel 0 = x, 1 = y (core/utils/utils.py:82-85)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack((xs, ys), dim=0).float()[None].expand(batch, -1, -1, -1)
