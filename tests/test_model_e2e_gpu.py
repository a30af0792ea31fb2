"""End-to-end criterion of BASELINE.json: final flow after 12 refinement iterations within 0.01 px mean EPE of the
reference path, identical random-init weights and synthetic frames -- checked on the test-side restatement of the
StreamFlow refinement loop (tests/model_harness.py), reference L1 operators (torch, same device, TF32 off) vs the
B200 operators."""
import copy

import pytest
import torch

from oracle import torch_port as tp
from tests import model_harness as mh

pytestmark = pytest.mark.gpu


class _A:
    pass


class RefAttention(torch.nn.Module):
    """The reference's Attention.forward (core/gma.py:53-65) on torch ops; same parameter names."""

    def __init__(self):
        super().__init__()
        self.to_qk = torch.nn.Conv2d(128, 256, 1, bias=False)

    def forward(self, fmap):
        return tp.cpu_attention(fmap, self.to_qk.weight.reshape(256, 128), heads=1, dim_head=128)


class RefAggregate(torch.nn.Module):
    """The reference's Aggregate.forward (core/gma.py:91-104) on torch ops; same parameter names."""

    def __init__(self):
        super().__init__()
        self.to_v = torch.nn.Conv2d(128, 128, 1, bias=False)
        self.gamma = torch.nn.Parameter(torch.zeros(1))

    def forward(self, attn, fmap):
        p, c, h, w = fmap.shape
        v = self.to_v(fmap).reshape(p, 1, c, h * w)
        out = torch.matmul(attn, v.transpose(2, 3)).transpose(2, 3).reshape(p, c, h, w)
        return fmap + self.gamma * out


def ref_corr(f1, f2, radius=4):
    return tp.CpuCorrPyramid(f1, f2, num_levels=4, radius=radius)


def build_pair(T, seed, flow_gain=0.01):
    import streamflow_b200 as sfb
    ref = mh.randomise(mh.FlowModel(ref_corr, RefAttention(), RefAggregate(), T=T), seed=seed,
                       flow_gain=flow_gain).cuda().eval()
    ours = mh.FlowModel(sfb.CorrBlock, sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128),
                        sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128), T=T).cuda().eval()
    missing, unexpected = ours.load_state_dict(copy.deepcopy(ref.state_dict()), strict=True)
    assert not missing and not unexpected          # same parameter names as the reference modules
    return ref, ours


def epe(a, b):
    return torch.sqrt(((a - b) ** 2).sum(dim=1)).mean().item()


@pytest.mark.parametrize("hw,T", [((184, 320), 4), ((440, 1024), 4)])
def test_final_flow_epe_vs_reference_path(hw, T):
    torch.backends.cuda.matmul.allow_tf32 = False      # the reference never enables TF32
    torch.backends.cudnn.allow_tf32 = False
    H, W = hw
    frames = mh.synthetic_clip(T, H, W, seed=1)
    ref, ours = build_pair(T, seed=3)
    up_r, low_r = ref(frames, iters=12)
    up_o, low_o = ours(frames, iters=12)
    torch.cuda.synchronize()
    worst = 0.0
    for i in range(T - 1):
        assert up_o[i].shape == (1, 2, H, W)
        mag = torch.sqrt((up_r[i] ** 2).sum(1)).mean().item()
        e = epe(up_o[i], up_r[i])
        worst = max(worst, e)
        assert torch.isfinite(up_o[i]).all()
        print(f"[{H}x{W}] pair {i}: |flow| {mag:.2f} px, mean EPE vs reference path {e:.5f} px")
        # Sintel-like magnitudes (a few px to ~15 px mean): the 0.01 px bound is absolute, so the random-init
        # network is scaled to produce realistic flow instead of running tens of pixels off
        assert 1.0 < mag < 20.0, f"unrealistic test: reference flow magnitude {mag:.3f} px"
        assert e < 0.01, f"pair {i}: mean EPE {e:.4f} px vs reference path (flow magnitude {mag:.2f} px)"
    print(f"[{H}x{W}] worst mean EPE {worst:.5f} px")


def test_gma_path_is_exercised():
    """gamma != 0 and a non-trivial temporal block: switching GMA off changes the flow (so the EPE test above
    really covers Attention / Aggregate)."""
    frames = mh.synthetic_clip(4, 184, 320, seed=1)
    ref, ours = build_pair(4, seed=3)
    up_a, _ = ours(frames, iters=4)
    with torch.no_grad():
        ours.update_block.aggregator.gamma.zero_()
    up_b, _ = ours(frames, iters=4)
    assert epe(up_a[0], up_b[0]) > 1e-3
