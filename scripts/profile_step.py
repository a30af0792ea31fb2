"""One Sintel-size hot-path step for ncu captures (keeps the kernel count small).

    ncu --set full --clock-control none --import-source on -k regex:<pattern> -c <n> -o gpurun_out/prof \
        python scripts/profile_step.py [iters] [reps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
hot = bench.HotPath(sfb, dev, host)
for rep in range(reps):
    feats, out = hot(t, iters)
torch.cuda.synchronize()
print("done")
