"""sf_pcblock_ffn1 at the Sintel size of convc1 (3 maps x 324 channels x 55 x 128): correctness + time against the eager
autocast ops of the reference (conv1x1, GELU, conv1x1, add + GELU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
import streamflow_b200 as sfb
torch.set_grad_enabled(False)
torch.manual_seed(0)
for C, P, h, w, dt in [(324, 3, 55, 128, torch.float32), (256, 3, 55, 128, torch.float16), (128, 3, 55, 128, torch.float16)]:
    H = int(1.5 * C)
    ffn1 = nn.Sequential(nn.Conv2d(C, H, 1), nn.GELU(), nn.Conv2d(H, C, 1)).cuda().eval()
    x = (torch.randn(P, C, h, w, device="cuda") * 0.7).to(dt)
    ref = F.gelu(x.float() + ffn1(x.float()))
    out = sfb.pcblock_ffn1(x, ffn1)
    torch.cuda.synchronize()
    err = ((out.float() - ref).norm() / ref.norm()).item()

    def t(fn, n=50):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n

    def eager():
        with torch.autocast("cuda", dtype=torch.float16):
            return F.gelu(x + ffn1(x))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sfb.pcblock_ffn1(x, ffn1); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                sfb.pcblock_ffn1(x, ffn1)
    us_graph = t(g.replay, 20) / 10
    flops = 2.0 * P * h * w * C * H * 2
    print(f"C={C} {dt}: rel err {err:.2e}; fused {t(lambda: sfb.pcblock_ffn1(x, ffn1)):.1f} us eager-launch, {us_graph:.1f} us back-to-back "
          f"({flops / us_graph / 1e6:.0f} TFLOP/s); reference autocast ops {t(eager):.1f} us")
