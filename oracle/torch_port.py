"""Torch restatement of the hot path in the reference's own op sequence (TEST / BASELINE INFRASTRUCTURE ONLY).

Runs on CPU (the timed baseline) and, being plain torch ops, on a CUDA device too -- the GPU tests use it as the
"reference L1 on the same device" arm of the end-to-end flow comparison (tests/test_model_e2e_gpu.py).

The reference's own implementation of this path is a sequence of PyTorch ops
(``matmul``, ``avg_pool2d``, ``grid_sample``, 1x1 ``conv2d``, ``softmax``); the
reference checkout does not travel to the GPU box, so this file restates that op
sequence for two purposes only:

  * ``bench.py``'s ``cpu_baseline`` object and ``bench.py --impl reference`` time it
    on the box's host cores (``kind: "port"``);
  * ``tests/`` cross-check it against the NumPy oracle and the committed golden vectors.

It is never imported by the product package ``streamflow_b200``.

Reference op sequence followed: core/corr.py:7-54, core/utils/utils.py:65-85,
core/gma.py:53-65 and :91-104.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def coords_grid(batch: int, ht: int, wd: int) -> torch.Tensor:
    """[B, 2, h, w] with channel 0 = x, 1 = y (core/utils/utils.py:82-85)."""
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack((xs, ys), 0).float()[None].repeat(batch, 1, 1, 1)


class CpuCorrPyramid:
    """Volume + pooled levels + 9x9 lookup on CPU, the reference's algorithm (core/corr.py)."""

    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4):
        b, d, h, w = fmap1.shape
        lhs = fmap1.reshape(b, d, h * w).transpose(1, 2)
        rhs = fmap2.reshape(b, d, h * w)
        vol = torch.matmul(lhs, rhs) / float(d) ** 0.5
        lvl = vol.reshape(b * h * w, 1, h, w)
        self.levels = [lvl]
        for _ in range(1, num_levels):
            lvl = F.avg_pool2d(lvl, 2, stride=2)
            self.levels.append(lvl)
        self.radius = radius
        off = torch.arange(-radius, radius + 1, dtype=torch.float32, device=fmap1.device)
        # window index i (slow) offsets x, j (fast) offsets y -- see SURVEY Appendix A.2
        self._dx = off.view(1, -1, 1).expand(1, off.numel(), off.numel())
        self._dy = off.view(1, 1, -1).expand(1, off.numel(), off.numel())

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        b, _, h, w = coords.shape
        cx = coords[:, 0].reshape(-1, 1, 1)
        cy = coords[:, 1].reshape(-1, 1, 1)
        chunks = []
        for lvl, img in enumerate(self.levels):
            hl, wl = img.shape[-2:]
            x = cx / 2 ** lvl + self._dx
            y = cy / 2 ** lvl + self._dy
            gx = 2 * x / (wl - 1) - 1
            gy = 2 * y / (hl - 1) - 1
            grid = torch.stack((gx, gy), dim=-1)
            s = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
            chunks.append(s.reshape(b, h, w, -1))
        return torch.cat(chunks, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def cpu_attention(fmap: torch.Tensor, w_qk: torch.Tensor, heads: int = 1, dim_head: int = 128) -> torch.Tensor:
    """softmax(scale * q k^T) over all positions (core/gma.py:53-65) -> [P, heads, N, N]."""
    p, c, h, w = fmap.shape
    qk = F.conv2d(fmap, w_qk.reshape(w_qk.shape[0], c, 1, 1).to(fmap.dtype))
    q, k = qk.chunk(2, dim=1)
    q = q.reshape(p, heads, dim_head, h * w) * dim_head ** -0.5
    k = k.reshape(p, heads, dim_head, h * w)
    sim = torch.matmul(q.transpose(2, 3), k)
    return sim.softmax(dim=-1)


def cpu_aggregate(attn: torch.Tensor, fmap: torch.Tensor, w_v: torch.Tensor, gamma: float,
                  w_proj: torch.Tensor | None = None, heads: int = 1) -> torch.Tensor:
    """fmap + gamma * (attn . to_v(fmap)) (core/gma.py:91-104)."""
    p, c, h, w = fmap.shape
    v = F.conv2d(fmap, w_v.reshape(w_v.shape[0], c, 1, 1).to(fmap.dtype))
    inner = v.shape[1]
    v = v.reshape(p, heads, inner // heads, h * w)
    out = torch.matmul(attn, v.transpose(2, 3))            # [P, heads, N, dh]
    out = out.transpose(2, 3).reshape(p, inner, h, w)
    if w_proj is not None:
        out = F.conv2d(out, w_proj.reshape(w_proj.shape[0], inner, 1, 1))
    return fmap + gamma * out


@torch.no_grad()
def cpu_hot_path(fmaps, coords_per_iter, inps, mfs, w_qk, w_v, gamma):
    """One clip through the hot path on CPU (same contract as streamflow_oracle.hot_path)."""
    t = fmaps.shape[1]
    pyrs = [CpuCorrPyramid(fmaps[:, i], fmaps[:, i + 1]) for i in range(t - 1)]
    attn = cpu_attention(inps, w_qk)
    feats = agg = None
    for it in range(coords_per_iter.shape[0]):
        feats = torch.stack([pyrs[i](coords_per_iter[it, i]) for i in range(t - 1)], 0)
        agg = cpu_aggregate(attn, mfs, w_v, gamma)
    return feats, agg
