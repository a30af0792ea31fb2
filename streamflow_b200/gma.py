"""Drop-in replacement for ``Attention`` / ``Aggregate`` of the reference's ``core/gma.py`` (sm_100a).

Same constructors, parameter names and call signatures as the reference (``core/gma.py:34-104``), so
reference checkpoints load unchanged (``to_qk.weight [2*inner, dim, 1, 1]``, ``to_v.weight [inner, dim, 1, 1]``,
``gamma [1]``) and ``core/models/streamflow.py:124`` / ``core/update.py:769`` run unmodified:

    attn = Attention(args=..., dim=128, heads=1, max_pos_size=160, dim_head=128)(inps)
    out  = Aggregate(args=..., dim=128, heads=1, dim_head=128)(attn, motion_features)

``Attention.forward`` returns an opaque ``AttentionHandle`` instead of the dense ``[P, 1, N, N]`` fp32 matrix
(the callers only pass it on to ``Aggregate``): it owns the fp16 softmax numerators E, their row sums and the
workspace.  ``handle.dense()`` materialises the reference-shaped matrix for tests.  ``Aggregate.forward`` streams the
motion features against E and applies ``to_v`` AFTER the attention-weighted sum inside the same kernel
(``sum_j (W_v x)_j a_ij = W_v (sum_j x_j a_ij)``), so no per-iteration projection launch exists.

The authors' memory-saving convention (``demo.py:235-282``, ``test_memory.py:240-282``) is accepted too:
``Attention(..., return_qk=True)`` returns the projected ``(q, k)`` like their ``Attention`` and
``Aggregate.forward(q, k, fmap)`` takes them in place of the handle.  Where the reference then redoes the whole
attention with ``flash_attn_func`` every refinement iteration, this module builds E once per distinct (q, k) pair
and reuses it for as long as the same, unmodified tensors come back.

The kernels are specialised for the shipped configuration heads=1, dim=dim_head=128; anything else raises
(there is no eager fallback).  Both modules are INFERENCE-ONLY: there are no backward kernels, so a call made with
autograd enabled on inputs or parameters that require grad raises ``StreamCorrError`` instead of silently returning
tensors without ``grad_fn`` (the parameters exist so that reference checkpoints load unchanged).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _lib
from ._lib import StreamCorrError
from .corr import _aligned_workspace, _on_device, _stream_ptr


class AttentionHandle:
    """What ``Attention.forward`` returns: E (tile-major fp16 softmax numerators, see include/streamcorr.h),
    rowsum [P, N] fp32 and the workspace."""

    def __init__(self, E, rowsum, ws_buf, ws_ptr, ws_bytes, shape):
        self.E, self.rowsum = E, rowsum
        self._ws_buf, self._ws_ptr, self._ws_bytes = ws_buf, ws_ptr, ws_bytes
        self.P, self.C, self.h, self.w = shape
        self.N = self.h * self.w

    @property
    def shape(self):
        return (self.P, 1, self.N, self.N)

    def dense(self):
        """Reference-shaped softmax matrix [P, 1, N, N] fp32 (test helper; O(N^2) memory)."""
        P, N = self.P, self.N
        mt, npad = (N + 127) // 128, self.E.numel() // (P * ((N + 127) // 128) * 128)
        # every 16 KB block is the 128B-swizzled image: logical 16-byte chunk c of row r sits at chunk c ^ (r & 7)
        blocks = self.E.view(P, mt, npad // 64, 128, 8, 8)
        r = torch.arange(128, device=self.E.device)[:, None]
        c = torch.arange(8, device=self.E.device)[None, :]
        idx = (c ^ (r & 7)).view(1, 1, 1, 128, 8, 1).expand_as(blocks)
        e = blocks.gather(4, idx).reshape(P, mt, npad // 64, 128, 64)
        e = e.permute(0, 1, 3, 2, 4).reshape(P, mt * 128, npad)
        return (e[:, :N, :N].float() / self.rowsum[:, :, None])[:, None]

    def row_sums_of_e(self):
        """sum_j E[p, i, j] in fp32 (test helper: must equal rowsum)."""
        P, N = self.P, self.N
        mt = (N + 127) // 128
        npad = self.E.numel() // (P * mt * 128)
        e = self.E.view(P, mt, npad // 64, 128, 64).float().sum(-1).sum(2)       # [P, mt, 128]
        return e.reshape(P, mt * 128)[:, :N]


def _version_of(t):
    """In-place modification counter, or None for inference tensors (they do not track versions)."""
    return None if t.is_inference() else t._version


def _weight_2d(module, attr, param, rows, cols, dtype=torch.float32):
    """contiguous [rows, cols] copy/view of a conv weight in `dtype`, cached until the parameter is modified in
    place or replaced (keeps the reshape / conversion off every call).  The cache holds the parameter tensor itself
    and compares identity, so a new tensor that happens to reuse the storage address cannot hit; inference tensors
    (no version counter) are never cached."""
    ver = _version_of(param)
    cached = getattr(module, attr, None)
    if (cached is None or ver is None or cached[0] is not param or
            cached[1] != (param.data_ptr(), ver, param.dtype, param.device)):
        w = param.detach().reshape(rows, cols)
        if w.dtype != dtype or not w.is_contiguous():
            w = w.to(dtype).contiguous()
        cached = (param, (param.data_ptr(), ver, param.dtype, param.device), w)
        object.__setattr__(module, attr, cached)
    return cached[2]


def _check_cfg(dim, heads, dim_head):
    if heads != 1 or dim != 128 or dim_head != 128:
        raise StreamCorrError(
            f"GMA kernels are specialised for heads=1, dim=dim_head=128 (got heads={heads}, dim={dim}, "
            f"dim_head={dim_head}); no fallback path exists")


_GMA_PRECISION = os.environ.get("STREAMCORR_GMA_PRECISION", "f16x2")


class Attention(nn.Module):
    """``precision`` (attribute, default from STREAMCORR_GMA_PRECISION, "f16x2"): "f16x2" keeps the q/k projections
    and logits fp32-faithful (hi/lo-split fp16 operands; measured 1.6e-4 on the attention matrix vs the fp32 oracle);
    "f16" rounds q, k to fp16 for the logit GEMM -- the operand precision of the reference's own autocast path, 7e-4
    vs the fp32 oracle, and 4 % faster on the whole step."""

    def __init__(self, *, args, dim, max_pos_size=100, heads=4, dim_head=128, return_qk=False):
        super().__init__()
        self.precision = _GMA_PRECISION
        self.return_qk = return_qk
        self.args = args
        self.heads = heads
        self.dim = dim
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        inner_dim = heads * dim_head
        self.to_qk = nn.Conv2d(dim, inner_dim * 2, 1, bias=False)

    def forward(self, fmap):
        _check_cfg(self.dim, self.heads, self.dim_head)
        if fmap.dim() != 4 or fmap.shape[1] != self.dim:
            raise StreamCorrError(f"Attention expects [P, {self.dim}, h, w], got {tuple(fmap.shape)}")
        if self.return_qk:
            # demo.py:268-273: the projection only; the attention itself happens inside Aggregate.forward(q, k, fmap)
            q, k = self.to_qk(fmap).chunk(2, dim=1)
            return q, k
        if self.precision not in ("f16", "f16x2"):
            raise StreamCorrError(f"unknown GMA precision {self.precision!r}; choose 'f16' or 'f16x2'")
        if not fmap.is_cuda:
            raise StreamCorrError("Attention needs a CUDA tensor (no CPU fallback)")
        _lib.require_no_grad("Attention", fmap, self.to_qk.weight)
        x = fmap.detach()
        if not x.is_contiguous():
            x = x.contiguous()
        P, C, h, w = x.shape
        N = h * w
        dev = x.device
        L = _lib.lib()
        wq = _weight_2d(self, "_wq_cache", self.to_qk.weight, 2 * self.dim_head, C)
        with _on_device(dev):
            npad = L.sf_gma_npad(N)
            E = torch.empty((L.sf_gma_e_elems(P, N),), dtype=torch.float16, device=dev)
            rowsum = torch.empty((P, N), dtype=torch.float32, device=dev)
            ws_bytes = L.sf_gma_workspace_bytes(P, C, N, self.dim_head)
            ws_buf, ws_ptr = _aligned_workspace(ws_bytes, dev)
            rc = L.sf_gma_attention(x.data_ptr(), _lib.torch_dtype_code(x.dtype), wq.data_ptr(), P, C, N,
                                    self.dim_head, float(self.scale), _lib.PRECISIONS[self.precision], E.data_ptr(),
                                    rowsum.data_ptr(), ws_ptr,
                                    ws_bytes, _stream_ptr(dev))
        _lib.check(rc, "sf_gma_attention")
        return AttentionHandle(E, rowsum, ws_buf, ws_ptr, ws_bytes, (P, C, h, w))


class Aggregate(nn.Module):
    def __init__(self, args, dim, heads=4, dim_head=128):
        super().__init__()
        self.args = args
        self.heads = heads
        self.dim = dim
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        inner_dim = heads * dim_head
        self.precision = _GMA_PRECISION          # used by the (q, k, fmap) convention only
        self.to_v = nn.Conv2d(dim, inner_dim, 1, bias=False)
        self.gamma = nn.Parameter(torch.zeros(1))
        if dim != inner_dim:
            self.project = nn.Conv2d(inner_dim, dim, 1, bias=False)
        else:
            self.project = None

    def _handle_from_qk(self, q, k):
        """E / rowsum for already-projected q, k [P, dim_head, h, w]; cached while the same tensors come back."""
        if q.shape != k.shape or q.dim() != 4 or q.shape[1] != self.dim_head:
            raise StreamCorrError(f"Aggregate(q, k, fmap) expects q, k of shape [P, {self.dim_head}, h, w], got "
                                  f"{tuple(q.shape)} and {tuple(k.shape)}")
        if not q.is_cuda or q.device != k.device or q.dtype != k.dtype:
            raise StreamCorrError("Aggregate(q, k, fmap) needs q and k on the same CUDA device with the same dtype")
        _lib.require_no_grad("Aggregate(q, k, fmap)", q, k)
        vq, vk = _version_of(q), _version_of(k)
        key = (q.data_ptr(), k.data_ptr(), vq, vk, tuple(q.shape), q.dtype, q.device, self.precision)
        cached = getattr(self, "_qk_cache", None)
        # inference tensors carry no version counter: an in-place update could not be detected, so never reuse
        if (cached is not None and vq is not None and vk is not None and cached[0] == key and
                cached[1][0] is q and cached[1][1] is k):
            return cached[2]
        qc, kc = q.detach().contiguous(), k.detach().contiguous()
        P, d, h, w = qc.shape
        N = h * w
        dev = qc.device
        L = _lib.lib()
        with _on_device(dev):
            E = torch.empty((L.sf_gma_e_elems(P, N),), dtype=torch.float16, device=dev)
            rowsum = torch.empty((P, N), dtype=torch.float32, device=dev)
            ws_bytes = L.sf_gma_workspace_bytes(P, d, N, d)
            ws_buf, ws_ptr = _aligned_workspace(ws_bytes, dev)
            rc = L.sf_gma_attention_qk(qc.data_ptr(), kc.data_ptr(), _lib.torch_dtype_code(qc.dtype), P, N, d,
                                       float(self.scale), _lib.PRECISIONS[self.precision], E.data_ptr(),
                                       rowsum.data_ptr(), ws_ptr, ws_bytes, _stream_ptr(dev))
        _lib.check(rc, "sf_gma_attention_qk")
        handle = AttentionHandle(E, rowsum, ws_buf, ws_ptr, ws_bytes, (P, d, h, w))
        # the cache holds q and k so their storage cannot be recycled under the key
        object.__setattr__(self, "_qk_cache", (key, (q, k), handle))
        return handle

    def forward(self, *inputs):
        """``forward(attn, fmap)`` (core/gma.py:91) or ``forward(querys, keys, fmap)`` (demo.py:237)."""
        _check_cfg(self.dim, self.heads, self.dim_head)
        if len(inputs) == 3:
            attn, fmap = self._handle_from_qk(inputs[0], inputs[1]), inputs[2]
        elif len(inputs) == 2:
            attn, fmap = inputs
        else:
            raise TypeError(f"Aggregate.forward takes (attn, fmap) or (querys, keys, fmap), got {len(inputs)} arguments")
        if not isinstance(attn, AttentionHandle):
            raise StreamCorrError("Aggregate expects the handle returned by streamflow_b200.gma.Attention")
        if fmap.dim() != 4 or tuple(fmap.shape) != (attn.P, self.dim, attn.h, attn.w):
            raise StreamCorrError(f"Aggregate expects fmap [{attn.P}, {self.dim}, {attn.h}, {attn.w}], "
                                  f"got {tuple(fmap.shape)}")
        _lib.require_no_grad("Aggregate", fmap, self.to_v.weight, self.gamma)
        x = fmap.detach()
        if not x.is_contiguous():
            x = x.contiguous()
        P, C, h, w = x.shape
        dev = x.device
        # W_v enters the in-kernel tensor-core GEMM as fp16 (as under the reference's autocast): convert once per weight update
        wv = _weight_2d(self, "_wv_cache", self.to_v.weight, self.dim_head, C, torch.float16)
        gamma = self.gamma
        if gamma.dtype != torch.float32:
            gamma = gamma.detach().float()
        with _on_device(dev):
            out = torch.empty((P, C, h, w), dtype=torch.float32, device=dev)
            rc = _lib.lib().sf_gma_aggregate(attn.E.data_ptr(), attn.rowsum.data_ptr(), x.data_ptr(),
                                             _lib.torch_dtype_code(x.dtype), wv.data_ptr(), _lib.DT_F16,
                                             gamma.data_ptr(),
                                             out.data_ptr(), P, C, attn.N, self.dim_head, attn._ws_ptr,
                                             attn._ws_bytes, _stream_ptr(dev))
        _lib.check(rc, "sf_gma_aggregate")
        return out
