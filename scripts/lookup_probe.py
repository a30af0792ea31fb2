"""Where does the lookup's time go?  Back-to-back launches inside one CUDA graph (Sintel size, 3 pairs) with
(a) the bench's random-walk coordinates, (b) the SAME coordinates every launch (maximal L2 reuse: the 57 MB of
window tiles fit the 126 MB L2), (c) one pair only, (d) far out-of-image coordinates (no window loads at all: the
store + fixed cost), (e) fp16 output.  Tells whether L2 hits can pay for this gather at all."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb

dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
fm = host["fm_nhwc"].to(dev).permute(0, 1, 4, 2, 3)
coords = host["coords"].to(dev)
group = sfb.CorrGroup.from_fmaps(fm, radius=4)
group16 = sfb.CorrGroup(group.blocks, out_dtype=torch.float16)
one = sfb.CorrGroup(group.blocks[:1])
far = coords + 4000.0


def graph_time(fn, calls=12, replays=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0); fn(1)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            keep = [fn(i) for i in range(calls)]
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * calls)


rows = [
    ("3 pairs, random-walk coords (bench)", lambda i: group([coords[i % 12, j] for j in range(3)])),
    ("3 pairs, SAME coords every launch", lambda i: group([coords[0, j] for j in range(3)])),
    ("1 pair, random-walk coords", lambda i: one([coords[i % 12, 0]])),
    ("1 pair, SAME coords every launch", lambda i: one([coords[0, 0]])),
    ("3 pairs, far outside the image (stores + fixed cost only)", lambda i: group([far[i % 12, j] for j in range(3)])),
    ("3 pairs, random-walk coords, fp16 output", lambda i: group16([coords[i % 12, j] for j in range(3)])),
]
for name, fn in rows:
    print(f"{graph_time(fn):8.2f} us  {name}")
