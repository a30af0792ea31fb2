"""First-contact diagnostics on the GPU box: prints error statistics per path instead of asserting."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import load_golden, rel_err, rs_normal
from oracle import streamflow_oracle as so
from streamflow_b200 import CorrBlock

def cuda(x): return torch.from_numpy(np.ascontiguousarray(x)).cuda()

print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
small = load_golden("corr_small.npz")
blk = CorrBlock.from_dense_pyramid([cuda(small[f"level{l}"]) for l in range(4)])
for name in ["grid", "half", "jitter", "far", "neg", "border"]:
    out = blk(cuda(small[f"coords_{name}"])).cpu().numpy()
    print(f"lookup[{name}] rel={rel_err(out, small[f'lookup_{name}']):.3e}")
for prec in ["fp32", "f16", "f16x2"]:
    try:
        b = CorrBlock(cuda(small["f1"]), cuda(small["f2"]), precision=prec)
        torch.cuda.synchronize()
        for l in range(4):
            got = b.corr_pyramid[l].cpu().numpy()
            print(f"build[{prec}] level{l} rel={rel_err(got, small[f'level{l}']):.3e} nan={np.isnan(got).sum()}")
    except Exception as e:
        print(f"build[{prec}] FAILED: {e}")
# bigger: cfg1 and sintel timing
for (h, w, d) in [(46, 62, 256), (55, 128, 256)]:
    f1 = cuda(rs_normal(0, (1, d, h, w))); f2 = cuda(rs_normal(100, (1, d, h, w)))
    ref = CorrBlock(f1, f2, precision="fp32"); torch.cuda.synchronize()
    for prec in ["fp32", "f16", "f16x2"]:
        try:
            for _ in range(2): b = CorrBlock(f1, f2, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): b = CorrBlock(f1, f2, precision=prec)
            e1.record(); torch.cuda.synchronize()
            errs = [float((b.corr_pyramid[l] - ref.corr_pyramid[l]).norm() / ref.corr_pyramid[l].norm()) for l in range(4)]
            print(f"{h}x{w} build[{prec}] {e0.elapsed_time(e1)/5*1e3:.1f} us  rel vs fp32-simt: " + " ".join(f"{e:.2e}" for e in errs))
        except Exception as e:
            print(f"{h}x{w} build[{prec}] FAILED: {e}")
    g = torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), 0).float()[None].cuda().contiguous()
    c = (g + 5 * torch.randn_like(g)).contiguous()
    for _ in range(3): o = ref(c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): o = ref(c)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{h}x{w} lookup {us:.1f} us  -> {2904*h*w/us/1e3:.1f} GB/s algorithmic (L2-warm, single pair)")

# ---------------------------------------------------------------- GMA
from streamflow_b200 import Attention, Aggregate
class _A: pass
g = load_golden("gma_small.npz")
att = Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
agg = Aggregate(args=_A(), dim=128, heads=1, dim_head=128).cuda()
with torch.no_grad():
    att.to_qk.weight.copy_(cuda(g["w_qk"]).view(256, 128, 1, 1)); agg.to_v.weight.copy_(cuda(g["w_v"]).view(128, 128, 1, 1)); agg.gamma.fill_(float(g["gamma"]))
try:
    h = att(cuda(g["inp"])); torch.cuda.synchronize()
    attn = h.dense().cpu().numpy()
    print(f"gma attn rel={rel_err(attn, g['attn']):.3e} nan={np.isnan(attn).sum()} rowsum_err={np.abs(attn.sum(-1)-1).max():.2e}")
    out = agg(h, cuda(g["mf"])).cpu().numpy()
    print(f"gma out rel={rel_err(out, g['out']):.3e} delta rel={rel_err(out-g['mf'], g['out']-g['mf']):.3e}")
except Exception as e:
    print("gma small FAILED:", e)
try:
    P, hh, ww = 3, 55, 128
    inp = torch.relu(torch.randn(P, 128, hh, ww, device="cuda")); mf = torch.randn(P, 128, hh, ww, device="cuda")
    for _ in range(2): h = att(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h = att(inp); e1.record(); torch.cuda.synchronize()
    print(f"sintel attention (once per clip): {e0.elapsed_time(e1)*1e3:.1f} us")
    for _ in range(3): o = agg(h, mf)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): o = agg(h, mf)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"sintel aggregate per iter: {us:.1f} us -> E stream {P*7040*7040*2/us/1e3:.0f} GB/s")
except Exception as e:
    print("gma sintel FAILED:", e)
