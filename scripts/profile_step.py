"""One Sintel-size hot-path step for ncu captures (keeps the kernel count small).

    ncu --set full --clock-control none --import-source on -k regex:<pattern> -c <n> -o gpurun_out/prof \
        python scripts/profile_step.py [iters] [reps] [extras]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
extras = len(sys.argv) > 3 and sys.argv[3] == "extras"      # capture these with -k regex:"lookup_reg|pcblock"
dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
hot = bench.HotPath(sfb, dev, host)
for rep in range(reps):
    feats, out = hot(t, iters)
# the kernels outside the 3-pair hot loop: the short-launch lookup (one pair per call, as the unmodified model calls it)
# and the motion encoder's entry on the lookup output
import torch.nn as nn
fm = t["fm_nhwc"].permute(0, 1, 4, 2, 3)
blk = sfb.CorrBlock(fm[:, 0], fm[:, 1], radius=4) if extras else None
ffn1 = nn.Sequential(nn.Conv2d(324, 486, 1), nn.GELU(), nn.Conv2d(486, 324, 1)).to(dev).eval()
for rep in range(reps if extras else 0):
    one = blk(t["coords"][0, 0])
    y = sfb.pcblock_ffn1(feats.reshape(3, 324, 55, 128), ffn1)
torch.cuda.synchronize()
print("done")
