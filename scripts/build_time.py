"""Group build (3 pairs, Sintel size) split into absmax + pack and the GEMM, per precision mode, back-to-back in a graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
import bench
import streamflow_b200 as sfb

dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
L = sfb.lib()
fm_exact = host["fm_nhwc"].to(dev).permute(0, 1, 4, 2, 3)
fm_rand = (host["fm_nhwc"].to(dev) * 1.0001).permute(0, 1, 4, 2, 3)        # no longer fp16-representable


def graph_time(fn, mask, calls=4, replays=5):
    L.sf_debug_select_kernels(7, mask)
    try:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn(); fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                keep = [fn() for _ in range(calls)]
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


for name, fm in (("fp16-exact inputs", fm_exact), ("arbitrary fp32 inputs", fm_rand)):
    for prec in ("f16", "auto", "f16x2"):
        f = lambda: sfb.CorrGroup.from_fmaps(fm, precision=prec)
        f()      # a full build first so the GEMM-only graph reads valid operands / flags
        print(f"{name:22s} {prec:6s}: absmax+pack {graph_time(f, 1):7.1f} us   gemm {graph_time(f, 2):7.1f} us   "
              f"build {graph_time(f, 3):7.1f} us")
