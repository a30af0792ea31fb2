"""Aggregate-only timing at Sintel size (graph of 12 back-to-back launches, alternating E buffers), for ablations via
STREAMCORR_AGG_DEBUG (bit 0 no rowsum, 1 no second GEMM, 2 no fmap tile fetch, 3 no L2 prefetch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
import bench
import streamflow_b200 as sfb

dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
class _A: pass
att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
att.to_qk.weight.copy_(host["w_qk"].view(256, 128, 1, 1)); agg.to_v.weight.copy_(host["w_v"].view(128, 128, 1, 1)); agg.gamma.fill_(0.8)
inps, mfs = host["inps"].to(dev), host["mfs"].to(dev)
handles = [att(inps), att(inps)]
L = sfb.lib()

def graph_time(mask, calls=12, replays=10):
    L.sf_debug_select_kernels(mask, 3)
    try:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            agg(handles[0], mfs); agg(handles[1], mfs)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                keep = [agg(handles[i & 1], mfs) for i in range(calls)]
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (replays * calls)
    finally:
        L.sf_debug_select_kernels(7, 3)

print(f"dbg={os.environ.get('STREAMCORR_AGG_DEBUG', '0'):>2}  cast {graph_time(1):6.2f} us   aggregate {graph_time(2):6.2f} us   both {graph_time(3):6.2f} us")
