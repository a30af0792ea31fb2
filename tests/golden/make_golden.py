"""Generate the golden vectors under tests/golden/ from the REFERENCE implementation.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference's own ``core/corr.py`` and ``core/gma.py`` (torch CPU, fp32,
``torch.no_grad``), feeds them inputs drawn from ``numpy.random.RandomState`` (a frozen
generator, so inputs can be regenerated anywhere without torch's RNG) and stores
inputs + outputs as ``.npz``.  The reference checkout does not travel to the GPU box;
these files do.  Nothing here is imported by the product package.

Cases
  corr_small.npz    B=1, D=32, 17x20 (odd dims -> floor-mode pooling); full pyramid +
                    lookups for six coordinate sets (integer grid, half-pixel, jitter,
                    far out-of-bounds, negative shift, border-straddling)
  corr_batch.npz    B=2, D=16, 16x16, channels-last *strided* inputs (as in the model,
                    core/models/streamflow.py:107,110)
  corr_cfg1.npz     BASELINE configs[0]: D=256, 46x62, seed 0 -- sampled known answers
                    (8192 random positions per tensor + sums), inputs regenerated from seed
  gma_small.npz     Attention + Aggregate, heads=1, dim=dim_head=128, 12x16, P=2
  gma_proj.npz      heads=2, dim_head=32, dim=96 (project branch, core/gma.py:86-89,99-100)
  upsample.npz      SKFlow_MF8.upsample_flow (core/models/streamflow.py:82-93), the reference METHOD itself, loaded
                    through oracle/ref_model.py (timm shim): N=2, 6x7 -> 48x56, fp32 masks
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, os.path.join(REF, "core"))
warnings.filterwarnings("ignore")
import corr as ref_corr  # noqa: E402
import gma as ref_gma  # noqa: E402
from utils.utils import coords_grid as ref_coords_grid  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def rs_normal(seed, shape):
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


def rs_uniform(seed, lo, hi, shape):
    return np.random.RandomState(seed).uniform(lo, hi, shape).astype(np.float32)


def coord_sets(seed, b, h, w):
    g = ref_coords_grid(b, h, w).numpy().copy()
    sets = {
        "grid": g,
        "half": g + 0.5,
        "jitter": g + 3.0 * rs_normal(seed + 10, g.shape),
        "far": g + rs_uniform(seed + 11, -80, 80, g.shape),
        "neg": g - 6.25,
        "border": g * np.float32(1.0) + rs_uniform(seed + 12, -1, 1, g.shape) * np.array(
            [w, h], np.float32).reshape(1, 2, 1, 1),
    }
    return {k: v.astype(np.float32) for k, v in sets.items()}


def sample_idx(seed, n, k=8192):
    return np.random.RandomState(seed).randint(0, n, size=min(k, n)).astype(np.int64)


def case_corr_small():
    f1, f2 = rs_normal(1, (1, 32, 17, 20)), rs_normal(2, (1, 32, 17, 20))
    cb = ref_corr.CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=4, radius=4)
    out = {"f1": f1, "f2": f2}
    for l, p in enumerate(cb.corr_pyramid):
        out[f"level{l}"] = p.numpy()
    for name, c in coord_sets(3, 1, 17, 20).items():
        out[f"coords_{name}"] = c
        out[f"lookup_{name}"] = cb(torch.from_numpy(c)).numpy()
    np.savez_compressed(os.path.join(OUT, "corr_small.npz"), **out)


def case_corr_batch():
    # channels-last storage viewed as [B, D, h, w]: exactly what fnet hands to CorrBlock
    a = rs_normal(4, (2, 16, 16, 16))  # [B, h, w, D]
    b = rs_normal(5, (2, 16, 16, 16))
    f1 = torch.from_numpy(a).permute(0, 3, 1, 2)
    f2 = torch.from_numpy(b).permute(0, 3, 1, 2)
    assert not f1.is_contiguous()
    cb = ref_corr.CorrBlock(f1, f2, radius=4)
    c = coord_sets(6, 2, 16, 16)["jitter"]
    out = {"f1_nhwc": a, "f2_nhwc": b, "coords": c, "lookup": cb(torch.from_numpy(c)).numpy()}
    for l, p in enumerate(cb.corr_pyramid):
        out[f"level{l}"] = p.numpy()
    np.savez_compressed(os.path.join(OUT, "corr_batch.npz"), **out)


def case_corr_cfg1():
    h, w, d = 46, 62, 256
    f1, f2 = rs_normal(0, (1, d, h, w)), rs_normal(100, (1, d, h, w))
    cb = ref_corr.CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), radius=4)
    out = {"shape": np.array([1, d, h, w]), "seed_f1": np.array(0), "seed_f2": np.array(100)}
    for l, p in enumerate(cb.corr_pyramid):
        flat = p.numpy().reshape(-1)
        idx = sample_idx(200 + l, flat.size)
        out[f"level{l}_idx"], out[f"level{l}_val"] = idx, flat[idx]
        out[f"level{l}_sum"] = np.array(flat.astype(np.float64).sum())
        out[f"level{l}_abs"] = np.array(np.abs(flat.astype(np.float64)).sum())
    for name, c in coord_sets(7, 1, h, w).items():
        flat = cb(torch.from_numpy(c)).numpy().reshape(-1)
        idx = sample_idx(300 + len(name), flat.size)
        out[f"lookup_{name}_idx"], out[f"lookup_{name}_val"] = idx, flat[idx]
        out[f"lookup_{name}_sum"] = np.array(flat.astype(np.float64).sum())
        out[f"lookup_{name}_abs"] = np.array(np.abs(flat.astype(np.float64)).sum())
    np.savez_compressed(os.path.join(OUT, "corr_cfg1.npz"), **out)


class _Args:
    pass


def case_gma(name, heads, dim_head, dim, h, w, p, seed):
    att = ref_gma.Attention(args=_Args(), dim=dim, heads=heads, max_pos_size=160, dim_head=dim_head)
    agg = ref_gma.Aggregate(args=_Args(), dim=dim, heads=heads, dim_head=dim_head)
    inner = heads * dim_head
    w_qk = rs_normal(seed, (2 * inner, dim)) * np.float32(dim ** -0.5) * np.float32(2.0)
    w_v = rs_normal(seed + 1, (inner, dim)) * np.float32(dim ** -0.5)
    gamma = np.float32(0.7)
    att.to_qk.weight.copy_(torch.from_numpy(w_qk).view(2 * inner, dim, 1, 1))
    agg.to_v.weight.copy_(torch.from_numpy(w_v).view(inner, dim, 1, 1))
    agg.gamma.fill_(float(gamma))
    out = {"w_qk": w_qk, "w_v": w_v, "gamma": np.array(gamma), "heads": np.array(heads),
           "dim_head": np.array(dim_head)}
    if agg.project is not None:
        w_p = rs_normal(seed + 2, (dim, inner)) * np.float32(inner ** -0.5)
        agg.project.weight.copy_(torch.from_numpy(w_p).view(dim, inner, 1, 1))
        out["w_proj"] = w_p
    inp = np.maximum(rs_normal(seed + 3, (p, dim, h, w)), 0)      # inps = relu(.) in the model
    mf = rs_normal(seed + 4, (p, dim, h, w))
    attn = att(torch.from_numpy(inp))
    res = agg(attn, torch.from_numpy(mf))
    out.update(inp=inp, mf=mf, attn=attn.numpy(), out=res.numpy())
    np.savez_compressed(os.path.join(OUT, name), **out)


def case_upsample():
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import ref_model as rm
    mod = rm.load_model_module("reference")
    flow = rs_normal(40, (2, 2, 6, 7)) * np.float32(4.0)
    mask = rs_normal(41, (2, 576, 6, 7)) * np.float32(2.0)
    up = mod.SKFlow_MF8.upsample_flow(None, torch.from_numpy(flow), torch.from_numpy(mask), ratio=8)
    np.savez_compressed(os.path.join(OUT, "upsample.npz"), flow=flow, mask=mask, out=up.numpy())


def case_pcblock():
    """The reference's own PCBlock4_Deep_nopool_res (core/update.py:12-36, imported through the timm shim like the model):
    inputs, ffn1 parameters, and the value of the first line of its forward, `F.gelu(x + self.ffn1(x))` (core/update.py:31),
    in fp32 on the CPU."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import ref_model as rm
    mod = rm.load_model_module("reference")
    upd = mod._l1_modules["update"]
    torch.manual_seed(50)
    blk = upd.PCBlock4_Deep_nopool_res(48, 32, [1, 15]).eval()
    x = torch.from_numpy(rs_normal(51, (2, 48, 5, 7)) * np.float32(1.5))
    with torch.no_grad():
        first = torch.nn.functional.gelu(x + blk.ffn1(x))
        full = blk(x)
    np.savez_compressed(os.path.join(OUT, "pcblock.npz"), x=x.numpy(), w1=blk.ffn1[0].weight.detach().numpy().reshape(72, 48),
                        b1=blk.ffn1[0].bias.detach().numpy(), w2=blk.ffn1[2].weight.detach().numpy().reshape(48, 72),
                        b2=blk.ffn1[2].bias.detach().numpy(), first=first.numpy(), block_out=full.numpy())


if __name__ == "__main__":
    case_upsample()
    case_pcblock()
    case_corr_small()
    case_corr_batch()
    case_corr_cfg1()
    case_gma("gma_small.npz", heads=1, dim_head=128, dim=128, h=12, w=16, p=2, seed=20)
    case_gma("gma_proj.npz", heads=2, dim_head=32, dim=96, h=6, w=10, p=1, seed=30)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
