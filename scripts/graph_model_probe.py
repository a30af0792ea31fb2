"""Whole-forward CUDA-graph replay of the unmodified reference model on the B200 operators (GraphedModel):
capture succeeds? flows identical to the eager forward? time per clip?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import streamflow_b200 as sfb  # noqa: E402
from streamflow_b200.graph import GraphedModel  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ref, ours = bench.load_full_models(dev)
H, W = 436, 1024
frames = bench.make_frames(4, H, W, 0).pin_memory()
runner = bench.FullModelRunner(ours, dev, H, W)


def barrier():
    torch.cuda.synchronize()


eager = runner.flows_on_device(frames).clone()
ms_eager = bench._time_events(lambda: runner(frames), 5, 2, barrier)
t0 = time.time()
gm = GraphedModel(ours, (4, 3, H, W), iters=12)
torch.cuda.synchronize()
print(f"capture took {time.time() - t0:.1f} s; {gm.launches} libstreamcorr launches inside the graph")
host_flows = torch.empty((3, 2, H, W)).pin_memory()


def step():
    host_flows.copy_(gm(frames), non_blocking=True)


g = gm(frames).clone()
epe = torch.sqrt(((g - eager) ** 2).sum(1)).mean(dim=(1, 2))
print("graph vs eager mean EPE per pair:", [float(x) for x in epe], "max abs diff", float((g - eager).abs().max()))
frames2 = bench.make_frames(4, H, W, 5).pin_memory()
e2 = runner.flows_on_device(frames2).clone()
g2 = gm(frames2).clone()
print("second clip: max abs diff", float((g2 - e2).abs().max()), "flow magnitude", float(e2.norm(dim=1).mean()))
ms_graph = bench._time_events(step, 10, 3, barrier)
print(f"eager {ms_eager:.2f} ms  graph {ms_graph:.2f} ms per clip  ({3 / ms_graph * 1e3:.1f} flows/s)")
print(f"memory: allocated {torch.cuda.memory_allocated() / 2**30:.2f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB")
