"""torch.profiler table of one full forward of the unmodified model on the B200 operators: where do the ~80 ms go?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ref, ours = bench.load_full_models(dev)
H, W = 436, 1024
frames = bench.make_frames(4, H, W, 0).pin_memory()
runner = bench.FullModelRunner(ours, dev, H, W)
for _ in range(2):
    runner(frames)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner(frames)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
