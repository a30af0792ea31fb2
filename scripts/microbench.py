"""Back-to-back timing of each hot-path op (no event gaps between launches): true kernel cost at Sintel size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb
from streamflow_b200 import _lib

dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
class _A: pass
att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
with torch.no_grad():
    att.to_qk.weight.copy_(host["w_qk"].view(256, 128, 1, 1)); agg.to_v.weight.copy_(host["w_v"].view(128, 128, 1, 1)); agg.gamma.fill_(0.8)
fmaps = t["fm_nhwc"].permute(0, 1, 4, 2, 3)
blocks = [sfb.CorrBlock(fmaps[:, i], fmaps[:, i + 1], radius=4) for i in range(3)]
group = sfb.CorrGroup(blocks)
handle = att(t["inps"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(name, fn, n=24, cold=False):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    if cold:
        tot = 0.0
        for i in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print(f"{name:28s} {tot / n * 1e3:8.1f} us  (L2 flushed before each launch, single launch incl. ~launch latency)")
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:28s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us  (back-to-back x{n})")

coords = [[t["coords"][it, i] for i in range(3)] for it in range(12)]


def graph_time(name, fn, n=12):
    """GPU-only cost per call: n calls captured in one CUDA graph (no CPU launch cost), replayed 5x."""
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
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1) / (5 * n) * 1e3:8.1f} us per call (graph of {n})")


graph_time("group lookup (3 pairs)", lambda i: group(coords[i % 12]))
graph_time("single-pair lookup", lambda i: blocks[0](coords[i % 12][0]))
graph_time("aggregate (proj + agg)", lambda i: agg(handle, t["mfs"]))
graph_time("corr build (1 pair)", lambda i: sfb.CorrBlock(fmaps[:, i % 3], fmaps[:, i % 3 + 1], radius=4), n=6)
graph_time("corr group build (3 pairs)", lambda i: sfb.CorrGroup.from_fmaps(fmaps, radius=4), n=4)
graph_time("corr build f16x2 (1 pair)", lambda i: sfb.CorrBlock(fmaps[:, i % 3], fmaps[:, i % 3 + 1], radius=4, precision="f16x2"), n=6)
graph_time("attention (3 maps)", lambda i: att(t["inps"]), n=3)

L = _lib.lib() if hasattr(_lib, "lib") else None
if L is not None:
    for mask, name in ((1, "  v projection only"), (2, "  aggregate kernel only")):
        L.sf_debug_select_kernels(mask, 3)
        try:
            graph_time(name, lambda i: agg(handle, t["mfs"]))
        finally:
            L.sf_debug_select_kernels(7, 3)
