"""CUDA-graph replay of a hot-path closure ("graph-captured iteration").

Every call of this package only enqueues kernels on the current stream (no host sync, no allocation inside the C
library), so a whole clip -- builds, attention, the 12 x (lookup, aggregate) loop -- can be captured once and
replayed with new data written into the same input buffers.  At Sintel size that removes ~45 kernel launches of
CPU-side latency per clip (1.53 ms eager -> 1.45 ms replayed on a B200).

    step = GraphedCall(lambda: hot_path(static_inputs))
    static_inputs["fmaps"].copy_(new_fmaps)           # refresh inputs in place
    feats, out = step()                                # replay; results live in the graph's static output tensors
"""
from __future__ import annotations

import torch

from . import _lib


class GraphedCall:
    """Capture ``fn()`` -- a closure over tensors whose storage stays put -- and replay it on the current stream.

    ``fn`` runs ``warmup`` times eagerly first (lazy initialisation such as ``cudaFuncSetAttribute`` must not happen
    during capture).  The tensors it returns are owned by the graph's memory pool and are overwritten by every
    replay: copy them out (or finish consuming them) before the next call.  ``launches`` is the number of kernels
    of this library inside one replay.
    """

    def __init__(self, fn, warmup: int = 2, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        L = _lib.lib()
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            before = L.sf_launch_count()
            with torch.cuda.graph(self.graph, stream=side):
                self.result = fn()
            self.launches = int(L.sf_launch_count() - before)
        torch.cuda.current_stream(dev).wait_stream(side)

    def __call__(self):
        self.graph.replay()
        return self.result
