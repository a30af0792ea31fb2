"""Wire / on-disk formats either side of the hot path, host-side only.

* ``InputPadder`` -- replicate-pad frames so H and W are multiples of 8 and crop flows back
  (``core/utils/utils.py:7-31``; 'sintel' mode pads symmetrically, any other mode pads the bottom only).
* ``write_flo`` / ``read_flo`` -- Middlebury ``.flo``: float32 magic 202021.25, int32 width, int32 height, then
  row-major interleaved (u, v) float32 (``core/utils/frame_utils.py:13-30, 86-115``).
* ``write_flow_kitti`` / ``read_flow_kitti`` -- KITTI 16-bit flow PNG (``core/utils/frame_utils.py:118-123, 137-141``).
  (``.flo5`` needs h5py, which is not part of this environment: not provided.)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

FLO_MAGIC = np.float32(202021.25)


class InputPadder:
    """Replicate-padding to a multiple of ``factor`` with the reference's API (``pad``, ``pad_list``, ``unpad``) and
    its split of the padding: ``mode="sintel"`` centres it (436 rows -> 2 above, 2 below), any other mode keeps the
    top edge and pads the bottom only (KITTI), columns are always centred.  ``_pad`` = [left, right, top, bottom]."""

    def __init__(self, dims, mode="sintel", factor=8):
        height, width = int(dims[-2]), int(dims[-1])
        self.ht, self.wd = height, width
        extra_h, extra_w = -height % factor, -width % factor
        left = extra_w // 2
        top = extra_h // 2 if mode == "sintel" else 0
        self._pad = [left, extra_w - left, top, extra_h - top]

    def _apply(self, x):
        return F.pad(x, self._pad, mode="replicate") if any(self._pad) else x

    def pad(self, *inputs):
        return [self._apply(x) for x in inputs]

    def pad_list(self, inputs):
        return [self._apply(x) for x in inputs]

    def unpad(self, x):
        left, right, top, bottom = self._pad
        return x[..., top:x.shape[-2] - bottom, left:x.shape[-1] - right]


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


# ---------------------------------------------------------------------------------------------------------
# KITTI flow PNG (``core/utils/frame_utils.py:118-123, 137-141``): 16-bit RGB, R = 64*u + 2^15, G = 64*v + 2^15,
# B = valid.  Written / read with zlib only (the reference goes through cv2, which stores its BGR array as RGB).
# ---------------------------------------------------------------------------------------------------------
def _png_chunk(tag: bytes, data: bytes) -> bytes:
    import struct
    import zlib
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_flow_kitti(path, flow, valid=None) -> None:
    """flow: [2, H, W] or [H, W, 2] (u, v) in pixels; valid: optional [H, W] mask (default all ones).
    Same quantisation as the reference's ``writeFlowKITTI``: ``uint16(64 * uv + 2**15)`` (truncation)."""
    import struct
    import zlib
    a = flow.detach().cpu().numpy() if isinstance(flow, torch.Tensor) else np.asarray(flow)
    if a.ndim != 3:
        raise ValueError(f"flow must be 3-D, got shape {a.shape}")
    if a.shape[0] == 2 and a.shape[2] != 2:
        a = np.transpose(a, (1, 2, 0))
    if a.shape[2] != 2:
        raise ValueError(f"flow needs 2 channels, got shape {a.shape}")
    h, w = a.shape[:2]
    q = 64.0 * a.astype(np.float32) + 2 ** 15                 # float32 arithmetic, as the reference does on float32 flows
    if q.min() < 0 or q.max() >= 65536:
        raise ValueError("flow outside the KITTI PNG range (-512, 512) px")
    v = np.ones((h, w), np.uint16) if valid is None else (np.asarray(valid).reshape(h, w) != 0).astype(np.uint16)
    rgb = np.concatenate([q.astype(np.uint16), v[..., None]], axis=-1)            # [H, W, 3] = (u, v, valid)
    raw = np.empty((h, 1 + w * 6), np.uint8)
    raw[:, 0] = 0                                                                 # filter type 0 (None) per scanline
    raw[:, 1:] = rgb.astype(">u2").view(np.uint8).reshape(h, w * 6)
    ihdr = struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)                          # 16-bit, colour type 2 (RGB)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", ihdr) + _png_chunk(b"IDAT", zlib.compress(raw.tobytes(), 6))
                + _png_chunk(b"IEND", b""))


def read_flow_kitti(path):
    """Returns (flow [H, W, 2] float32, valid [H, W] float32) like the reference's ``readFlowKITTI``."""
    import struct
    import zlib
    data = open(path, "rb").read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, w = 8, [], None
    while pos < len(data):
        n, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if depth != 16 or ctype != 2 or interlace != 0:
                raise ValueError(f"{path}: expected a non-interlaced 16-bit RGB PNG (KITTI flow)")
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"IEND":
            break
        pos += 12 + n
    if w is None:
        raise ValueError(f"{path}: missing IHDR")
    stride, bpp = w * 6, 6
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(h, 1 + stride)
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):                                          # undo the per-scanline PNG filters (types 0-4)
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros(stride, np.int32)
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                elif ft == 4:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError(f"{path}: bad PNG filter type {ft}")
                cur[i] = (line[i] + pred) & 255
        out[y] = cur.astype(np.uint8)
        prev = cur
    rgb = out.view(">u2").reshape(h, w, 3).astype(np.float32)
    return (rgb[:, :, :2] - 2 ** 15) / 64.0, rgb[:, :, 2]
