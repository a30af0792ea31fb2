"""CPU oracle for the StreamFlow correlation / GMA hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-formula NumPy restatement of the reference operators.  It is
the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``streamflow_b200``) never routes through it and fails loudly when the
CUDA library is missing.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports
``/root/reference/core/{corr,gma}.py`` (this container only) and commits seeded
input/output vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function below against them.

Reference semantics followed (paths relative to the reference checkout):
  corr_volume       core/corr.py:46-54   (fmap1^T . fmap2 / sqrt(D))
  build_pyramid     core/corr.py:7-21    (reshape to [B*N,1,h,w]; 3x avg_pool2d(2,2), floor mode)
  lookup            core/corr.py:23-44 + core/utils/utils.py:65-79
                    (x-major 9x9 window, grid_sample bilinear / zeros / align_corners=True)
  coords_grid       core/utils/utils.py:82-85  (channel 0 = x, channel 1 = y)
  attention         core/gma.py:53-65    (1x1 to_qk, scale q, softmax over all positions)
  aggregate         core/gma.py:91-104   (1x1 to_v, attn.v, optional project, fmap + gamma*out)

Everything is vectorised NumPy; ``dtype`` selects float32 (the reference's
arithmetic type) or float64 (a higher-precision cross-check of the oracle itself).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "corr_volume", "avg_pool2x2", "build_pyramid", "coords_grid", "lookup",
    "attention", "aggregate", "hot_path",
]


# --------------------------------------------------------------------------- corr
def corr_volume(fmap1: np.ndarray, fmap2: np.ndarray, dtype=np.float32) -> np.ndarray:
    """All-pairs correlation, core/corr.py:46-54.

    fmap1, fmap2: [B, D, h, w].  Returns [B, h, w, 1, h, w] with
    ``out[b,y,x,0,v,u] = sum_k f1[b,k,y,x] * f2[b,k,v,u] / sqrt(D)``.
    """
    f1 = np.asarray(fmap1, dtype=dtype)
    f2 = np.asarray(fmap2, dtype=dtype)
    if f1.shape != f2.shape or f1.ndim != 4:
        raise ValueError("fmap1/fmap2 must be 4-D with identical shapes")
    B, D, h, w = f1.shape
    a = f1.reshape(B, D, h * w)
    b = f2.reshape(B, D, h * w)
    vol = np.matmul(a.transpose(0, 2, 1), b)
    vol = vol / np.sqrt(dtype(D))
    return vol.reshape(B, h, w, 1, h, w).astype(dtype, copy=False)


def avg_pool2x2(x: np.ndarray) -> np.ndarray:
    """F.avg_pool2d(x, 2, stride=2): floor mode, trailing odd row/col dropped (core/corr.py:20)."""
    n, c, h, w = x.shape
    h2, w2 = h // 2, w // 2
    v = x[:, :, : 2 * h2, : 2 * w2].reshape(n, c, h2, 2, w2, 2)
    s = (v[:, :, :, 0, :, 0] + v[:, :, :, 0, :, 1]) + (v[:, :, :, 1, :, 0] + v[:, :, :, 1, :, 1])
    return (s * x.dtype.type(0.25)).astype(x.dtype, copy=False)


def build_pyramid(fmap1, fmap2, num_levels: int = 4, dtype=np.float32):
    """CorrBlock.__init__, core/corr.py:7-21.  Returns a list of [B*N, 1, h_l, w_l]."""
    vol = corr_volume(fmap1, fmap2, dtype)
    B, h, w = vol.shape[0], vol.shape[1], vol.shape[2]
    lvl = vol.reshape(B * h * w, 1, h, w)
    pyr = [lvl]
    for _ in range(num_levels - 1):
        lvl = avg_pool2x2(lvl)
        pyr.append(lvl)
    return pyr


def coords_grid(batch: int, ht: int, wd: int, dtype=np.float32) -> np.ndarray:
    """core/utils/utils.py:82-85: [B, 2, h, w], channel 0 = x (column), channel 1 = y (row)."""
    ys, xs = np.meshgrid(np.arange(ht), np.arange(wd), indexing="ij")
    g = np.stack([xs, ys], axis=0).astype(dtype)
    return np.broadcast_to(g[None], (batch, 2, ht, wd)).copy()


def _bilinear_zeros(img: np.ndarray, ix: np.ndarray, iy: np.ndarray) -> np.ndarray:
    """grid_sample(bilinear, zeros) on unnormalised pixel coords.

    img: [M, H, W]; ix, iy: [M, ...] pixel coordinates.  Per-tap zero padding.
    """
    M, H, W = img.shape
    dt = img.dtype.type
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    ax = (ix - x0).astype(img.dtype)
    ay = (iy - y0).astype(img.dtype)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    midx = np.arange(M).reshape((M,) + (1,) * (ix.ndim - 1))

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[midx, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok, v, dt(0))

    one = dt(1)
    out = (one - ax) * (one - ay) * tap(y0, x0)
    out = out + ax * (one - ay) * tap(y0, x0 + 1)
    out = out + (one - ax) * ay * tap(y0 + 1, x0)
    out = out + ax * ay * tap(y0 + 1, x0 + 1)
    return out.astype(img.dtype, copy=False)


def lookup(pyramid, coords: np.ndarray, radius: int = 4, exact_roundtrip: bool = True) -> np.ndarray:
    """CorrBlock.__call__, core/corr.py:23-44.

    pyramid: list of [B*N, 1, h_l, w_l]; coords: [B, 2, h, w] (ch0 = x).  Returns
    [B, L*(2r+1)^2, h, w]; channel ``l*81 + i*9 + j`` samples level l at
    ``(x, y) = (cx/2^l + i - r, cy/2^l + j - r)`` -- the FIRST window index moves x.

    ``exact_roundtrip`` reproduces the reference's normalise (utils.py:69-70) ->
    grid_sample de-normalise fp32 round trip; with False the pixel coordinate is used
    directly (differs by a few ulp of the coordinate; bilinear is continuous).
    """
    dt = pyramid[0].dtype.type
    coords = np.asarray(coords, dtype=pyramid[0].dtype)
    B, two, h, w = coords.shape
    assert two == 2
    cx = coords[:, 0].reshape(-1)
    cy = coords[:, 1].reshape(-1)
    r = radius
    d = np.arange(-r, r + 1).astype(pyramid[0].dtype)
    feats = []
    for lvl, vol in enumerate(pyramid):
        img = vol[:, 0]
        H, W = img.shape[1:]
        X = cx[:, None, None] / dt(2 ** lvl) + d[None, :, None]     # i -> x offset
        Y = cy[:, None, None] / dt(2 ** lvl) + d[None, None, :]     # j -> y offset
        X = np.broadcast_to(X, (X.shape[0], 2 * r + 1, 2 * r + 1)).astype(img.dtype)
        Y = np.broadcast_to(Y, (Y.shape[0], 2 * r + 1, 2 * r + 1)).astype(img.dtype)
        if exact_roundtrip:
            xn = dt(2) * X / dt(W - 1) - dt(1)
            yn = dt(2) * Y / dt(H - 1) - dt(1)
            X = ((xn + dt(1)) / dt(2)) * dt(W - 1)
            Y = ((yn + dt(1)) / dt(2)) * dt(H - 1)
        s = _bilinear_zeros(img, X, Y)                                # [BN, 9, 9]
        feats.append(s.reshape(B, h, w, (2 * r + 1) ** 2))
    out = np.concatenate(feats, axis=-1)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2)).astype(np.float32, copy=False)


# ---------------------------------------------------------------------------- gma
def _conv1x1(w: np.ndarray, x: np.ndarray) -> np.ndarray:
    """nn.Conv2d(kernel 1, bias=False): w [O, I] (or [O, I, 1, 1]), x [P, I, h, w]."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    return np.einsum("oi,pihw->pohw", w2, x, optimize=True)


def attention(fmap: np.ndarray, w_qk: np.ndarray, heads: int = 1, dim_head: int = 128,
              dtype=np.float32) -> np.ndarray:
    """gma.Attention.forward, core/gma.py:53-65.  Returns attn [P, heads, N, N]."""
    x = np.asarray(fmap, dtype=dtype)
    w = np.asarray(w_qk, dtype=dtype)
    P, C, h, wd = x.shape
    inner = heads * dim_head
    qk = _conv1x1(w, x)
    q, k = qk[:, :inner], qk[:, inner:]
    q = q.reshape(P, heads, dim_head, h * wd) * dtype(dim_head ** -0.5)
    k = k.reshape(P, heads, dim_head, h * wd)
    sim = np.einsum("phdi,phdj->phij", q, k, optimize=True)
    sim = sim - sim.max(axis=-1, keepdims=True)
    e = np.exp(sim)
    return (e / e.sum(axis=-1, keepdims=True)).astype(dtype, copy=False)


def aggregate(attn: np.ndarray, fmap: np.ndarray, w_v: np.ndarray, gamma: float,
              w_proj: np.ndarray | None = None, heads: int = 1, dtype=np.float32) -> np.ndarray:
    """gma.Aggregate.forward, core/gma.py:91-104.  Returns fmap + gamma * (attn . v)."""
    x = np.asarray(fmap, dtype=dtype)
    P, C, h, wd = x.shape
    v = _conv1x1(np.asarray(w_v, dtype=dtype), x)
    inner = v.shape[1]
    dh = inner // heads
    v = v.reshape(P, heads, dh, h * wd)
    out = np.einsum("phij,phdj->phdi", np.asarray(attn, dtype=dtype), v, optimize=True)
    out = out.reshape(P, inner, h, wd)
    if w_proj is not None:
        out = _conv1x1(np.asarray(w_proj, dtype=dtype), out)
    return (x + dtype(gamma) * out).astype(dtype, copy=False)


# ----------------------------------------------------------------- motion-encoder entry
def _erf(x: np.ndarray) -> np.ndarray:
    """erf by its Maclaurin / continued-fraction free closed form is not in NumPy: use math.erf element-wise (exact to double
    rounding; the fixtures are small)."""
    import math
    return np.vectorize(math.erf, otypes=[np.float64])(x)


def gelu(x: np.ndarray) -> np.ndarray:
    """torch.nn.GELU() / F.gelu default (approximate='none'): 0.5 x (1 + erf(x / sqrt 2))."""
    x = np.asarray(x, dtype=np.float64)
    return 0.5 * x * (1.0 + _erf(x / np.sqrt(2.0)))


def pcblock_ffn1(x: np.ndarray, w1: np.ndarray, b1: np.ndarray, w2: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """First line of PCBlock4_Deep_nopool_res.forward, core/update.py:31, with ffn1 = Conv2d(C, 1.5 C, 1) -> GELU ->
    Conv2d(1.5 C, C, 1) (core/update.py:18-22):  gelu(x + W2 . gelu(W1 . x + b1) + b2)  per pixel.  x: [P, C, h, w];
    w1: [H, C], w2: [C, H].  float64 throughout."""
    x = np.asarray(x, dtype=np.float64)
    hid = gelu(np.einsum("oc,pchw->pohw", np.asarray(w1, np.float64), x) + np.asarray(b1, np.float64)[None, :, None, None])
    y = np.einsum("oc,pchw->pohw", np.asarray(w2, np.float64), hid) + np.asarray(b2, np.float64)[None, :, None, None]
    return gelu(x + y)


# ----------------------------------------------------------------- whole hot path
def hot_path(fmaps, coords_per_iter, inps, mfs, w_qk, w_v, gamma, dtype=np.float32):
    """One clip through the hot path, the shape ``bench.py`` times.

    fmaps [B, T, D, h, w]; coords_per_iter [iters, T-1, B, 2, h, w]; inps [B*(T-1), d, h, w];
    mfs [B*(T-1), d, h, w] (motion features fed to Aggregate each iteration).
    Mirrors core/models/streamflow.py:110,123-124,132 + core/update.py:769.
    Returns (corr features of the last iteration [T-1, B, 324, h, w], aggregate output).
    """
    B, T = fmaps.shape[:2]
    pyrs = [build_pyramid(fmaps[:, i], fmaps[:, i + 1], dtype=dtype) for i in range(T - 1)]
    attn = attention(inps, w_qk, dtype=dtype)
    feats = agg = None
    for it in range(coords_per_iter.shape[0]):
        feats = np.stack([lookup(pyrs[i], coords_per_iter[it, i]) for i in range(T - 1)], 0)
        agg = aggregate(attn, mfs, w_v, gamma, dtype=dtype)
    return feats, agg
