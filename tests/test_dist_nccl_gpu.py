"""The one collective of the design on hardware: streaming windows / clips sharded over 2 GPUs and the NCCL gather of
the output flows (dist.run_windows / run_clips / gather_flows; demo.py:518-532 semantics).  Skips below 2 GPUs."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streamflow_b200 import dist as sfd

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs for NCCL")]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import streamflow_b200 as sfb
        n_frames, T, H, W = 19, 4, 32, 48                     # 6 windows: 3 + 3, no tail; then a ragged tail case
        frames = [torch.full((3, H, W), float(i), device=dev) for i in range(n_frames)]

        def flow_fn(win):
            # flow of pair (a, b): channel 0 = a, channel 1 = b, run through the upsample kernel's device (CUDA op)
            return [torch.stack([win[k][0], win[k + 1][0]]) for k in range(T - 1)]

        tm = {}
        flows = sfd.run_windows(frames, flow_fn, T=T, timings=tm)
        ok = flows.is_cuda and flows.shape == (n_frames - 1, 2, H, W)
        ok &= all(float(flows[i, 0, 0, 0]) == i and float(flows[i, 1, 0, 0]) == i + 1 for i in range(n_frames - 1))
        ok &= "gather_events" in tm and tm["gather_bytes"] == world * max(9, 9) * 2 * H * W * 4
        # ragged: 12 frames -> windows at 0, 3, 6 and a re-anchored tail at 8 that keeps 2 of its 3 flows
        flows = sfd.run_windows(frames[:12], flow_fn, T=T)
        ok &= flows.shape == (11, 2, H, W) and [float(f[0, 0, 0]) for f in flows] == [float(i) for i in range(11)]
        # clips: 5 clips over 2 ranks (3 + 2), a real kernel of the library on each rank's device
        clips = [torch.full((T - 1, 2, 4, 6), float(c), device=dev) for c in range(5)]
        masks = torch.zeros(T - 1, 576, 4, 6, device=dev)
        out = sfd.run_clips(clips, lambda c: sfb.upsample_flow(c, masks))
        ok &= out.shape == (5 * (T - 1), 2, 32, 48)
        ok &= [round(float(out[i * (T - 1), 0, 12, 12]) / 8.0, 4) for i in range(5)] == [float(c) for c in range(5)]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_window_sharding_and_flow_gather_over_nccl():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(0, True), (1, True)]
