"""GPU-only time of one hot-path step: capture the step in a CUDA graph and replay it (no CPU launch cost)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb

dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
class _A: pass
att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
with torch.no_grad():
    att.to_qk.weight.copy_(host["w_qk"].view(256, 128, 1, 1)); agg.to_v.weight.copy_(host["w_v"].view(128, 128, 1, 1)); agg.gamma.fill_(0.8)
fmaps = t["fm_nhwc"].permute(0, 1, 4, 2, 3)

def step():
    group = sfb.CorrGroup.from_fmaps(fmaps, radius=4)
    handle = att(t["inps"])
    for it in range(12):
        feats = group([t["coords"][it, i] for i in range(3)])
        out = agg(handle, t["mfs"])
    return feats, out

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"eager: CPU enqueue {1e3*(t1-t0):.3f} ms, until GPU done {1e3*(t2-t0):.3f} ms")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        res = step()
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/10:.3f} ms per step (GPU only)")
