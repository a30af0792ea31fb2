"""Parity at the sizes of BASELINE.json configs[2] (KITTI 376x1248, batched clips) and configs[3] (Spring 1080x1920)
against the reference's torch op sequence run on the same GPU (oracle/torch_port.py, fp32, TF32 off), plus the
streaming-window driver of configs[4] on one rank.  Tolerance: 1e-3 norm-wise (north_star)."""
import pytest
import torch

import streamflow_b200
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class _A:
    pass


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.cuda.empty_cache()


def test_kitti_batched_clips():
    """47x156 (N = 7332, w/4 = 39 tiles; pooled widths 78/39/19 are ragged), B = 2 clips in one CorrBlock."""
    from streamflow_b200 import CorrBlock, CorrGroup, coords_grid
    torch.manual_seed(0)
    B, h, w, d = 2, 47, 156, 256
    fm = torch.randn(B, 3, h, w, d, device="cuda").half().float().permute(0, 1, 4, 2, 3)
    ours = [CorrBlock(fm[:, i], fm[:, i + 1]) for i in range(2)]
    refs = [tp.CpuCorrPyramid(fm[:, i], fm[:, i + 1]) for i in range(2)]
    for l in range(4):
        assert rel(ours[0].corr_pyramid[l], refs[0].levels[l]) < 1e-3
    coords = [coords_grid(B, h, w, device="cuda").contiguous() + 4.0 * torch.randn(B, 2, h, w, device="cuda")
              for _ in range(2)]
    for i in range(2):
        assert rel(ours[i](coords[i]), refs[i](coords[i])) < 1e-3
    # pair-batched launch returns the (B T) C H W tensor of streamflow.py:132, rows ordered b*(T-1) + t
    grouped = CorrGroup(ours)(coords)
    stacked = torch.stack([refs[i](coords[i]) for i in range(2)], dim=1).flatten(0, 1)
    assert grouped.shape == stacked.shape == (B * 2, 324, h, w)
    assert rel(grouped, stacked) < 1e-3
    half = CorrGroup(ours, out_dtype=torch.float16)(coords)
    assert half.dtype == torch.float16 and rel(half.float(), stacked) < 2e-3


def test_spring_size_pair():
    """135x240 -> N = 32400: 4.2 GB level 0 per pair, 5.6 GB pyramid; GMA numerators 2.1 GB per map."""
    from streamflow_b200 import Aggregate, Attention, CorrBlock, coords_grid
    torch.manual_seed(1)
    h, w, d = 135, 240, 256
    fm = torch.randn(1, 2, h, w, d, device="cuda").half().float().permute(0, 1, 4, 2, 3)
    blk = CorrBlock(fm[:, 0], fm[:, 1])
    ref = tp.CpuCorrPyramid(fm[:, 0], fm[:, 1])
    c = coords_grid(1, h, w, device="cuda").contiguous() + 6.0 * torch.randn(1, 2, h, w, device="cuda")
    out, want = blk(c), ref(c)
    assert out.shape == (1, 324, h, w)
    assert rel(out, want) < 1e-3
    # far out-of-bounds queries give exact zeros
    far = c + 4000.0
    assert float(blk(far).abs().max()) == 0.0
    del blk, ref, out, want
    torch.cuda.empty_cache()

    att = Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg = Aggregate(args=_A(), dim=128, heads=1, dim_head=128).cuda()
    with torch.no_grad():
        att.to_qk.weight.normal_(0, 0.15)
        agg.to_v.weight.normal_(0, 0.09)
        agg.gamma.fill_(1.1)
    inp = torch.relu(torch.randn(1, 128, h, w, device="cuda"))
    mf = torch.randn(1, 128, h, w, device="cuda")
    hd = att(inp)
    got = agg(hd, mf)
    attn = tp.cpu_attention(inp, att.to_qk.weight.detach().view(256, 128))
    want = tp.cpu_aggregate(attn, mf, agg.to_v.weight.detach().view(128, 128), 1.1)
    assert rel(got - mf, want - mf) < 1e-3


def test_streaming_windows_single_rank():
    """configs[4] driver on one rank: 10 frames -> 3 windows -> 9 flows in temporal order (demo.py:515-532)."""
    from streamflow_b200 import dist as sfd
    frames = [torch.full((3, 16, 24), float(i), device="cuda") for i in range(10)]

    def fake_model(win):      # flow k of a window encodes its first frame index
        return [torch.zeros(2, 16, 24, device="cuda") + float(win[k][0, 0, 0]) for k in range(3)]

    flows = sfd.run_windows(frames, fake_model, T=4)
    assert flows.shape == (9, 2, 16, 24)
    assert [int(f[0, 0, 0]) for f in flows] == list(range(9))


def test_kitti_gma_24_maps():
    """configs[2] on one GPU: 8 clips x 3 pairs = 24 attention maps at 47x156 (2.6 GB of fp16 numerators)."""
    from streamflow_b200 import Aggregate, Attention
    torch.manual_seed(2)
    P, h, w = 24, 47, 156
    att = Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg = Aggregate(args=_A(), dim=128, heads=1, dim_head=128).cuda()
    with torch.no_grad():
        att.to_qk.weight.normal_(0, 0.15)
        agg.to_v.weight.normal_(0, 0.09)
        agg.gamma.fill_(0.9)
    inp = torch.relu(torch.randn(P, 128, h, w, device="cuda"))
    mf = torch.randn(P, 128, h, w, device="cuda")
    hd = att(inp)
    got = agg(hd, mf)
    for p0 in (0, 11, 23):      # reference on single maps keeps the fp32 N x N matrix small
        attn = tp.cpu_attention(inp[p0:p0 + 1], att.to_qk.weight.detach().view(256, 128))
        want = tp.cpu_aggregate(attn, mf[p0:p0 + 1], agg.to_v.weight.detach().view(128, 128), 0.9)
        assert rel(got[p0:p0 + 1] - mf[p0:p0 + 1], want - mf[p0:p0 + 1]) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_aggregate_low_precision_motion_features(dtype):
    """Aggregate accepts fp16 / bf16 motion features (templated projection and epilogue); result is fp32."""
    from streamflow_b200 import Aggregate, Attention
    torch.manual_seed(3)
    P, h, w = 2, 24, 40
    att = Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg = Aggregate(args=_A(), dim=128, heads=1, dim_head=128).cuda()
    with torch.no_grad():
        att.to_qk.weight.normal_(0, 0.15)
        agg.to_v.weight.normal_(0, 0.09)
        agg.gamma.fill_(1.3)
    inp = torch.relu(torch.randn(P, 128, h, w, device="cuda"))
    mf = torch.randn(P, 128, h, w, device="cuda").to(dtype)
    got = agg(att(inp), mf)
    assert got.dtype == torch.float32
    attn = tp.cpu_attention(inp, att.to_qk.weight.detach().view(256, 128))
    want = tp.cpu_aggregate(attn, mf.float(), agg.to_v.weight.detach().view(128, 128), 1.3)
    assert rel(got - mf.float(), want - mf.float()) < 1e-3


def test_whole_step_is_cuda_graph_capturable():
    """Every entry point only enqueues work (no host sync, no allocation inside the library): a whole hot-path step
    can be captured in a CUDA graph and replayed with new inputs in the same buffers."""
    import streamflow_b200 as sfb
    torch.manual_seed(4)
    h, w = 24, 32
    fm = torch.randn(1, 3, h, w, 64, device="cuda").half().float().permute(0, 1, 4, 2, 3)
    inp = torch.relu(torch.randn(2, 128, h, w, device="cuda"))
    mf = torch.randn(2, 128, h, w, device="cuda")
    coords = [sfb.coords_grid(1, h, w, device="cuda").contiguous() + torch.randn(1, 2, h, w, device="cuda") for _ in range(2)]
    att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).cuda()
    with torch.no_grad():
        att.to_qk.weight.normal_(0, 0.15)
        agg.to_v.weight.normal_(0, 0.09)
        agg.gamma.fill_(0.7)

    def step():
        blocks = [sfb.CorrBlock(fm[:, i], fm[:, i + 1]) for i in range(2)]
        feats = sfb.CorrGroup(blocks)(coords)
        return feats, agg(att(inp), mf)

    eager = [t.clone() for t in step()]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            out = step()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], eager[0])
    assert torch.equal(out[1], eager[1])         # the aggregate has no atomics: bit-identical run to run
    mf.add_(1.0)                                 # new inputs in the same buffers
    g.replay()
    torch.cuda.synchronize()
    want = step()[1]
    assert torch.equal(out[1], want)


def test_graphed_call_helper_replays_the_step():
    """streamflow_b200.GraphedCall: capture once, refresh inputs in place, replay; counts the library's launches."""
    import streamflow_b200 as sfb
    torch.manual_seed(9)
    h, w = 16, 24
    fm = torch.randn(1, 4, h, w, 64, device="cuda").permute(0, 1, 4, 2, 3)
    inp = torch.relu(torch.randn(3, 128, h, w, device="cuda"))
    mf = torch.randn(3, 128, h, w, device="cuda")
    coords = [(sfb.coords_grid(1, h, w, device="cuda") + torch.randn(1, 2, h, w, device="cuda")).contiguous()
              for _ in range(3)]
    att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).cuda()
    with torch.no_grad():
        att.to_qk.weight.normal_(0, 0.15)
        agg.to_v.weight.normal_(0, 0.09)
        agg.gamma.fill_(0.9)

    def step():
        group = sfb.CorrGroup.from_fmaps(fm)
        handle = att(inp)
        out = None
        for _ in range(2):
            feats = group(coords)
            out = agg(handle, mf)
        return feats, out

    graphed = sfb.GraphedCall(step)
    L = sfb.lib()
    before = L.sf_launch_count()
    step()
    eager_launches = L.sf_launch_count() - before
    assert graphed.launches == eager_launches and 12 <= eager_launches <= 16, (graphed.launches, eager_launches)
    got = [t.clone() for t in graphed()]
    want = step()
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    coords[0].add_(0.5)
    mf.mul_(2.0)
    got = graphed()
    torch.cuda.synchronize()
    want = step()
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


@pytest.mark.parametrize("B", [1, 2])
def test_group_build_from_fmaps_matches_per_pair_blocks(B):
    """CorrGroup.from_fmaps (one batched build of the T-1 pyramids) == T-1 CorrBlock builds,
    on channels-last strided frame features as the model produces them."""
    from streamflow_b200 import CorrBlock, CorrGroup
    T, D, h, w = 4, 256, 24, 40
    g = torch.Generator(device="cpu").manual_seed(5)
    fm = torch.randn(B, T, h, w, D, generator=g).cuda().permute(0, 1, 4, 2, 3)       # channels-last views
    coords = [(streamflow_b200.coords_grid(B, h, w, device="cuda") + 4 * torch.randn(B, 2, h, w, generator=g).cuda())
              .contiguous() for _ in range(T - 1)]
    group = CorrGroup.from_fmaps(fm, radius=4)
    blocks = [CorrBlock(fm[:, i], fm[:, i + 1], radius=4) for i in range(T - 1)]
    assert len(group.blocks) == T - 1
    for gb, blk in zip(group.blocks, blocks):
        for a, b in zip(gb.corr_pyramid, blk.corr_pyramid):
            assert a.shape == b.shape
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    out = group(coords)
    ref = torch.stack([blk(c) for blk, c in zip(blocks, coords)], dim=1).reshape(B * (T - 1), 324, h, w)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(streamflow_b200.StreamCorrError):
        CorrGroup.from_fmaps(fm[:, :1])


def test_sintel_size_direct_parity_vs_reference_ops():
    """BASELINE configs[1] itself -- 55x128, D=256, the 3 pairs of a T=4 clip, the bench's own inputs -- compared
    DIRECTLY (not through properties) with the reference op sequence on the same GPU (fp32, TF32 off): full pyramids,
    lookup features of a random-walk iteration, attention + aggregation."""
    import bench
    import streamflow_b200 as sfb
    from oracle import torch_port as tp
    torch.backends.cuda.matmul.allow_tf32 = False
    host = bench.make_inputs(0)
    dev = torch.device("cuda")
    fm = host["fm_nhwc"].to(dev).permute(0, 1, 4, 2, 3)
    coords = host["coords"].to(dev)
    group = sfb.CorrGroup.from_fmaps(fm, radius=4)
    feats = group([coords[5, i] for i in range(3)])
    for i in range(3):
        ref = tp.CpuCorrPyramid(fm[:, i], fm[:, i + 1])
        for l in range(4):
            got = group.blocks[i].corr_pyramid[l]
            err = float((got - ref.levels[l]).norm() / ref.levels[l].norm())
            assert err < 5e-6, f"pair {i} level {l}: {err:.3e}"          # fp16-exact inputs: exact products
        want = ref(coords[5, i])
        err = float((feats[i] - want[0]).norm() / want.norm())
        assert err < 1e-5, f"pair {i} lookup: {err:.3e}"
        del ref, want
    hot = bench.HotPath(sfb, dev, host)
    inps, mfs = host["inps"].to(dev), host["mfs"].to(dev)
    out = hot.agg(hot.att(inps), mfs)
    ref = tp.cpu_aggregate(tp.cpu_attention(inps, host["w_qk"].to(dev)), mfs, host["w_v"].to(dev), host["gamma"])
    err = float(((out - mfs) - (ref - mfs)).norm() / (ref - mfs).norm())
    assert err < 1e-3, f"gamma * attn . v at Sintel size: {err:.3e}"


def test_hot_path_is_bit_reproducible():
    """No atomics on floats anywhere on the path (integer row sums, one accumulator per output, fixed tile-to-CTA maps):
    two runs of the whole hot path on the same inputs agree bit for bit -- pair GEMM, lookup, attention, aggregate."""
    import bench
    import streamflow_b200 as sfb
    host = bench.make_inputs(3, 24, 40)
    dev = torch.device("cuda")
    t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
    hot = bench.HotPath(sfb, dev, host)
    f1, o1 = hot(t, 3)
    f2, o2 = hot(t, 3)
    torch.cuda.synchronize()
    assert torch.equal(f1, f2) and torch.equal(o1, o2)
