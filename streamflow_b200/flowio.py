"""Wire / on-disk formats either side of the hot path, host-side only.

* ``InputPadder`` -- replicate-pad frames so H and W are multiples of 8 and crop flows back
  (``core/utils/utils.py:7-31``; 'sintel' mode pads symmetrically, any other mode pads the bottom only).
* ``write_flo`` / ``read_flo`` -- Middlebury ``.flo``: float32 magic 202021.25, int32 width, int32 height, then
  row-major interleaved (u, v) float32 (``core/utils/frame_utils.py:13-30, 86-115``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

FLO_MAGIC = np.float32(202021.25)


class InputPadder:
    def __init__(self, dims, mode="sintel", factor=8):
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // factor) + 1) * factor - self.ht) % factor
        pad_wd = (((self.wd // factor) + 1) * factor - self.wd) % factor
        if mode == "sintel":
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def pad_list(self, inputs):
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        return x[..., self._pad[2]:ht - self._pad[3], self._pad[0]:wd - self._pad[1]]


def write_flo(path, flow) -> None:
    """flow: [2, H, W] tensor/array (u, v) or [H, W, 2]."""
    a = flow.detach().cpu().numpy() if isinstance(flow, torch.Tensor) else np.asarray(flow)
    if a.ndim != 3:
        raise ValueError(f"flow must be 3-D, got shape {a.shape}")
    if a.shape[0] == 2 and a.shape[2] != 2:
        a = np.transpose(a, (1, 2, 0))
    if a.shape[2] != 2:
        raise ValueError(f"flow needs 2 channels, got shape {a.shape}")
    h, w = a.shape[:2]
    with open(path, "wb") as f:
        FLO_MAGIC.tofile(f)
        np.int32(w).tofile(f)
        np.int32(h).tofile(f)
        np.ascontiguousarray(a, dtype=np.float32).tofile(f)


def read_flo(path) -> np.ndarray:
    """Returns [H, W, 2] float32."""
    with open(path, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != FLO_MAGIC:
            raise ValueError(f"{path}: not a Middlebury .flo file")
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
    if data.size != 2 * w * h:
        raise ValueError(f"{path}: truncated .flo file")
    return data.reshape(h, w, 2)
