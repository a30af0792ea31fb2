"""Acceptance test of BASELINE.json's north_star: the UNMODIFIED reference model -- SKFlow_MF8
(core/models/streamflow.py:30-147), SKUpdateBlock_TAM_v3 (core/update.py:739-782) and the Twins_CSC encoder
(core/encoders/twins_csc.py), executed from oracle/_ref -- runs on the B200 operators through
streamflow_b200.install(), and its final flow after 12 iterations stays within 0.01 px mean EPE of the same
model, same weights, same frames on the reference's own core/corr.py + core/gma.py (same GPU, TF32 off)."""
import os
import sys
import warnings

import pytest
import torch

from oracle import ref_model as rm
from tests import model_harness as mh

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rm.available(), reason="oracle/_ref not built (python oracle/make_ref.py)")]


def _pair(T, seed, mixed_precision=True, flow_gain=1.0):
    import streamflow_b200 as sfb
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_mod, our_mod = rm.load_model_module("reference"), rm.load_model_module("b200")
        torch.manual_seed(seed)
        ref = rm.randomise(rm.build_model(ref_mod, T, mixed_precision), seed=seed + 1, flow_gain=flow_gain).cuda().eval()
        ours = rm.build_model(our_mod, T, mixed_precision).cuda().eval()
    res = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert ours.update_block.aggregator.__class__ is sfb.Aggregate
    assert ours.att.__class__ is sfb.Attention and our_mod.CorrBlock is sfb.CorrBlock
    return ref, ours, ref_mod


def _epe(a, b):
    return torch.sqrt(((a - b) ** 2).sum(dim=1)).mean().item()


def _run(model, frames, iters):
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(frames, iters=iters, test_mode=True)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("hw,mixed", [((192, 320), True), ((436, 1024), True), ((192, 320), False)])
def test_unmodified_reference_model_epe(hw, mixed):
    torch.backends.cuda.matmul.allow_tf32 = False      # the reference never enables TF32
    torch.backends.cudnn.allow_tf32 = False
    H, W = hw
    T = 4
    ref, ours, ref_mod = _pair(T, seed=3, mixed_precision=mixed)
    core = os.path.dirname(os.path.dirname(ref_mod.__file__))
    sys.path.insert(0, core)
    try:
        from utils.utils import InputPadder          # the reference's own padder (core/utils/utils.py:7-31)
    finally:
        sys.path.remove(core)
    frames = mh.synthetic_clip(T, H, W, seed=1)
    padder = InputPadder(frames[0].shape)
    frames = padder.pad_list(frames) if hasattr(padder, "pad_list") else list(padder.pad(*frames))
    up_r = _run(ref, frames, 12)
    import streamflow_b200 as sfb
    n0 = sfb.lib().sf_launch_count()
    up_o = _run(ours, frames, 12)
    launched = sfb.lib().sf_launch_count() - n0
    assert launched >= 12 * (T - 1), f"only {launched} libstreamcorr launches: the B200 operators did not run"
    worst = 0.0
    for i in range(T - 1):
        fr, fo = padder.unpad(up_r[i]), padder.unpad(up_o[i])
        assert fo.shape == (1, 2, H, W) and torch.isfinite(fo).all()
        mag = torch.sqrt((fr ** 2).sum(1)).mean().item()
        e = _epe(fo, fr)
        worst = max(worst, e)
        print(f"[{H}x{W} amp={mixed}] pair {i}: |flow| {mag:.2f} px, mean EPE vs reference L1 {e:.5f} px")
        assert 0.5 < mag < 40.0, f"unrealistic test: reference flow magnitude {mag:.3f} px"
        assert e <= 0.01, f"pair {i}: mean EPE {e:.4f} px (bound 0.01 px; flow magnitude {mag:.2f} px)"
    print(f"[{H}x{W} amp={mixed}] worst mean EPE {worst:.5f} px, {launched} libstreamcorr launches")


def test_gma_is_live_in_the_reference_model():
    """gamma != 0 and the re-randomised temporal block: zeroing gamma changes the flow, so the EPE test covers GMA."""
    ref, ours, _ = _pair(4, seed=3)
    frames = mh.synthetic_clip(4, 192, 320, seed=1)
    a = _run(ours, frames, 4)
    with torch.no_grad():
        ours.update_block.aggregator.gamma.zero_()
    b = _run(ours, frames, 4)
    assert _epe(a[0], b[0]) > 1e-3
