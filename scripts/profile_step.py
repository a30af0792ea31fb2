"""One Sintel-size hot-path step for ncu captures (keeps the kernel count small).

    ncu --set full --clock-control none --import-source on -k regex:<pattern> -c <n> -o gpurun_out/prof \
        python scripts/profile_step.py [iters]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
class _A: pass
att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
with torch.no_grad():
    att.to_qk.weight.copy_(host["w_qk"].view(256, 128, 1, 1)); agg.to_v.weight.copy_(host["w_v"].view(128, 128, 1, 1)); agg.gamma.fill_(0.8)
fmaps = t["fm_nhwc"].permute(0, 1, 4, 2, 3)
for rep in range(2):
    group = sfb.CorrGroup.from_fmaps(fmaps, radius=4)
    handle = att(t["inps"])
    for it in range(iters):
        feats = group([t["coords"][it, i] for i in range(3)])
        out = agg(handle, t["mfs"])
torch.cuda.synchronize()
print("done")
