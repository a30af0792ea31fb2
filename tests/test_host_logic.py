"""Host-side logic: reference-import shim, streaming-window scheduler (demo.py:515-532), rank partitioning and
the flow gather over a world_size-2 gloo group (the N>1 path of bench.py / dist.py on CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streamflow_b200 import dist as sfd


def reference_schedule(n, T):
    """Literal restatement of the loop at demo.py:515-532 (flags per frame; flow k kept iff flags[k] != -1)."""
    out, i = [], 0
    while True:
        if i + T <= n:
            start, flags = i, list(range(i, i + T))
        else:
            start, flags = n - T, [-1 if j < i else j for j in range(n - T, n)]
        out.append((start, tuple(f != -1 for f in flags[: T - 1])))
        if i + T >= n:
            break
        i += T - 1
    return out


@pytest.mark.parametrize("n,T", [(64, 4), (10, 4), (9, 4), (4, 4), (5, 4), (7, 3), (23, 5)])
def test_window_schedule_matches_reference_loop(n, T):
    got = [(w.start, w.keep) for w in sfd.window_schedule(n, T)]
    assert got == reference_schedule(n, T)
    pairs = [w.start + k for w in sfd.window_schedule(n, T) for k, keep in enumerate(w.keep) if keep]
    assert pairs == list(range(n - 1)), "every consecutive pair exactly once, in order"


def test_config5_streaming_shape():
    """BASELINE configs[4]: 64 frames -> 21 windows / 63 flows, block-partitioned 3,3,3,3,3,2,2,2 over 8 ranks."""
    sched = sfd.window_schedule(64, 4)
    assert len(sched) == 21 and [w.start for w in sched][:3] == [0, 3, 6] and sched[-1].start == 60
    assert [len(r) for r in sfd.partition(21, 8)] == [3, 3, 3, 3, 3, 2, 2, 2]
    with pytest.raises(ValueError):
        sfd.window_schedule(3, 4)


def test_install_shim_resolves_reference_imports():
    import streamflow_b200 as sfb
    sfb.install()
    try:
        import corr
        import gma
        from gma import Aggregate, Attention, SpatioTemporalAggregate, TemporalAggregate  # noqa: F401  (update.py:4)
        assert corr.CorrBlock is sfb.CorrBlock and gma.Attention is sfb.Attention
        with pytest.raises(NotImplementedError):
            gma.TMMAggregate()
    finally:
        sfb.uninstall()
    assert "corr" not in sys.modules or sys.modules["corr"].__doc__ != "streamflow_b200 shim for the reference's core/corr.py"


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_frames, T, H, W = 12, 4, 6, 8
        frames = [torch.full((3, H, W), float(i)) for i in range(n_frames)]

        def flow_fn(win):
            # synthetic "model": flow of pair (a, b) encodes the frame indices so ordering can be checked
            return [torch.stack([torch.full((H, W), float(win[k][0, 0, 0])),
                                 torch.full((H, W), float(win[k + 1][0, 0, 0]))]) for k in range(T - 1)]

        flows = sfd.run_windows(frames, flow_fn, T=T)
        ok = flows.shape == (n_frames - 1, 2, H, W)
        ok &= all(float(flows[i, 0, 0, 0]) == i and float(flows[i, 1, 0, 0]) == i + 1 for i in range(n_frames - 1))
        clips = [torch.full((T - 1, 2, H, W), float(c)) for c in range(5)]
        out = sfd.run_clips(clips, lambda c: c + 0.5)
        ok &= out.shape == (5 * (T - 1), 2, H, W)
        ok &= [float(out[i * (T - 1), 0, 0, 0]) for i in range(5)] == [c + 0.5 for c in range(5)]
        # fewer clips than ranks: EVERY rank raises before any collective (no peer is left hanging in a gather)
        try:
            sfd.run_clips(clips[:1], lambda c: c)
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_windows_and_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_process_paths():
    frames = [torch.full((3, 4, 4), float(i)) for i in range(9)]
    flows = sfd.run_windows(frames, lambda w: [torch.zeros(2, 4, 4) + float(w[k][0, 0, 0]) for k in range(3)], T=4)
    assert flows.shape == (8, 2, 4, 4) and [float(f[0, 0, 0]) for f in flows] == list(range(8))


def test_next_row_entry_points_fail_loudly_without_a_device():
    """pcblock_ffn1 / GraphedModel / patch_motion_encoder host logic: CPU tensors and unsupported widths raise StreamCorrError
    before any kernel is touched (no CPU fallback), and the patch is reversible."""
    import torch
    import torch.nn as nn
    import streamflow_b200 as sfb
    from streamflow_b200 import pcblock

    ffn1 = nn.Sequential(nn.Conv2d(32, 48, 1), nn.GELU(), nn.Conv2d(48, 32, 1)).eval()
    with torch.no_grad():
        with pytest.raises(sfb.StreamCorrError):
            sfb.pcblock_ffn1(torch.randn(1, 32, 4, 4), ffn1)                 # not a CUDA tensor
    # weight packing: zero padding to (ceil128(hidden), ceil64(C)) / (ceil16(C), ceil128(hidden)), cached per module
    w1p, b1p, w2p, b2, C, Hd = pcblock._pack(ffn1)
    assert (C, Hd) == (32, 48) and w1p.shape == (128, 64) and w2p.shape == (32, 128) and b1p.shape == (128,)
    assert torch.equal(w1p[:48, :32].float(), ffn1[0].weight.detach().reshape(48, 32).half().float())
    assert float(w1p[48:].abs().max()) == 0.0 and float(w1p[:, 32:].abs().max()) == 0.0 and float(b1p[48:].abs().max()) == 0.0
    assert pcblock._pack(ffn1)[0] is w1p                                       # cache hit
    with torch.no_grad():
        ffn1[0].weight.mul_(2.0)
    assert pcblock._pack(ffn1)[0] is not w1p                                   # in-place update -> repacked
    with pytest.raises(sfb.StreamCorrError):
        pcblock._pack(nn.Sequential(nn.Conv2d(640, 960, 1), nn.GELU(), nn.Conv2d(960, 640, 1)))    # gru width: unsupported
    with pytest.raises(sfb.StreamCorrError):
        pcblock._pack(nn.Sequential(nn.Conv2d(32, 48, 3, padding=1), nn.GELU(), nn.Conv2d(48, 32, 1)))

    class Blk(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.ffn1 = nn.Sequential(nn.Conv2d(c, int(1.5 * c), 1), nn.GELU(), nn.Conv2d(int(1.5 * c), c, 1))
            self.conv_list = nn.ModuleList()
            self.pw = nn.Conv2d(c, c, 1)
            self.ffn2 = nn.Identity()

        def forward(self, x):
            return "reference forward"

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.update_block = nn.Module()
            self.update_block.encoder = nn.Module()
            self.update_block.encoder.convc1 = Blk(324)
            self.update_block.encoder.conv = Blk(640)                          # too wide: left alone

    m = Model().eval()
    assert sfb.patch_motion_encoder(m) == ["convc1"]
    assert "forward" in m.update_block.encoder.convc1.__dict__ and "forward" not in m.update_block.encoder.conv.__dict__
    sfb.unpatch_motion_encoder(m)
    assert m.update_block.encoder.convc1(None) == "reference forward"
    with pytest.raises(sfb.StreamCorrError):
        sfb.GraphedModel(m, (4, 3, 64, 64))                                    # model not on a CUDA device
