"""Back-to-back (CUDA-graph) op times at the other BASELINE.json config sizes: KITTI 376x1248 and Spring 1080x1920."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import streamflow_b200 as sfb

dev = torch.device("cuda", 0)
class _A: pass


def graph_time(fn, n):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(2): fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            keep = [fn(i) for i in range(n)]
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    del keep, g
    return e0.elapsed_time(e1) / (3 * n) * 1e3


for name, (h, w), T in (("Sintel 436x1024", (55, 128), 4), ("KITTI 376x1248", (47, 156), 4), ("Spring 1080x1920", (135, 240), 4)):
    torch.manual_seed(0)
    N, P = h * w, T - 1
    fm = torch.randn(1, T, h, w, 256, device=dev).half().float().permute(0, 1, 4, 2, 3)
    inps = torch.relu(torch.randn(P, 128, h, w, device=dev))
    mfs = torch.randn(P, 128, h, w, device=dev)
    coords = [(sfb.coords_grid(1, h, w, device=dev) + 5 * torch.randn(1, 2, h, w, device=dev)).contiguous() for _ in range(P)]
    att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
    agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
    with torch.no_grad():
        att.to_qk.weight.normal_(0, 0.1); agg.to_v.weight.normal_(0, 0.09); agg.gamma.fill_(0.8)
    n_small = 2 if N > 20000 else 4
    t_build = graph_time(lambda i: sfb.CorrGroup.from_fmaps(fm), n_small)
    group = sfb.CorrGroup.from_fmaps(fm)
    t_look = graph_time(lambda i: group(coords), 6)
    t_att = graph_time(lambda i: att(inps), n_small)
    handle = att(inps)
    t_agg = graph_time(lambda i: agg(handle, mfs), 6)
    pyr = sum(t.numel() * 4 for b in group.blocks for t in b._levels) / 1e9
    e_gb = handle.E.numel() * 2 / 1e9
    step = t_build + t_att + 12 * (t_look + t_agg)
    print(f"{name:18s} N={N:6d}  build(3 pairs) {t_build:9.1f} us  lookup(3) {t_look:8.1f} us  attention {t_att:9.1f} us  "
          f"aggregate {t_agg:8.1f} us  -> clip {step / 1e3:7.2f} ms  | pyramid {pyr:5.2f} GB, E {e_gb:5.2f} GB")
    del group, handle
    torch.cuda.empty_cache()
