"""Blind-debug aid for the aggregate kernel: compares against torch on the same GPU at several sizes and prints
where the error sits (per query block / per channel block), so one gpurun round-trip localises a bug."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
import streamflow_b200 as sfb

class _A: pass
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (P, h, w, dt) in [(2, 12, 16, torch.float32), (1, 24, 32, torch.float32), (3, 55, 128, torch.float32),
                      (3, 47, 156, torch.float32), (2, 37, 53, torch.float16), (24, 47, 156, torch.float32),
                      (1, 100, 160, torch.float32)]:
    N = h * w
    att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
    agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
    agg.gamma.fill_(0.8)
    inp = torch.relu(torch.randn(P, 128, h, w, device=dev))
    mf = torch.randn(P, 128, h, w, device=dev).to(dt)
    handle = att(inp)
    out = agg(handle, mf)
    torch.cuda.synchronize()
    # reference from the handle's own softmax (isolates the aggregate): fp64 on GPU in chunks
    wv = agg.to_v.weight.reshape(128, 128).double()
    ref = torch.empty_like(out, dtype=torch.float64)
    for pb in range(P):
        x = mf[pb].reshape(128, N).double()
        v = wv @ x                                     # [128, N]
        if N <= 8000:
            a = handle.dense()[pb, 0].double() if P * N * N < 3e8 else None
        else:
            a = None
        if a is None:
            hd = sfb.gma.AttentionHandle
            q, k = (att.to_qk.weight.reshape(256, 128).double() @ inp[pb].reshape(128, N).double()).chunk(2, 0)
            a = torch.softmax((q.t() * 128 ** -0.5) @ k, dim=-1)
        ref[pb] = (mf[pb].double().reshape(128, N) + 0.8 * (v @ a.t())).reshape(128, h, w)
    d = (out.double() - ref)
    rel = float(d.norm() / (ref - mf.double()).norm())
    print(f"P={P} {h}x{w} {dt}: rel err of gamma*attn*v = {rel:.3e}  max|d| = {float(d.abs().max()):.3e} finite={bool(torch.isfinite(out).all())}")
    if rel > 2e-3:
        dq = d.abs().amax(dim=(0, 1)).reshape(-1)          # per query
        blocks = dq.reshape(-1)[: N // 16 * 16].reshape(-1, 16).amax(1)
        bad = (blocks > 1e-3).nonzero().flatten().tolist()
        print("   bad 16-query units:", bad[:40], "of", blocks.numel())
        dc = d.abs().amax(dim=(0, 2, 3))
        print("   per-channel-octet max err:", [f"{float(x):.2e}" for x in dc.reshape(16, 8).amax(1)])
        print("   per-map max err:", [f"{float(x):.2e}" for x in d.abs().amax(dim=(1, 2, 3))])
