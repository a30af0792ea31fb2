"""the motion encoder's entry `x = F.gelu(x + self.ffn1(x))` (core/update.py:31) as one tcgen05 kernel.
Compared with the same torch ops in fp32 (TF32 off), with the reference's own PCBlock4_Deep_nopool_res under autocast
(oracle/_ref), and inside the unmodified model (flow EPE)."""
import warnings

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ffn1(C, seed, device="cuda"):
    torch.manual_seed(seed)
    H = int(1.5 * C)
    m = nn.Sequential(nn.Conv2d(C, H, 1), nn.GELU(), nn.Conv2d(H, C, 1)).to(device).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("C,P,h,w,dtype", [(324, 3, 55, 128, torch.float32), (324, 1, 17, 20, torch.float32),
                                            (256, 3, 47, 156, torch.float16), (128, 2, 24, 40, torch.float16),
                                            (64, 1, 9, 13, torch.float32), (336, 1, 16, 16, torch.float32)])
def test_ffn1_against_torch_fp32(C, P, h, w, dtype):
    import streamflow_b200 as sfb
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ffn1 = _ffn1(C, seed=C + h)
    x = (torch.randn(P, C, h, w, device="cuda") * 0.7).to(dtype)
    with torch.no_grad():
        ref = F.gelu(x.float() + ffn1(x.float()))
        out = sfb.pcblock_ffn1(x, ffn1)
    torch.cuda.synchronize()
    assert out.shape == x.shape and out.dtype == dtype and out.is_contiguous()
    e = _rel(out.float(), ref)
    # fp16 operands (x, W1, GELU(hidden), W2) with fp32 accumulation, as the reference's autocast path; fp16 output adds 5e-4
    assert e < 2e-3, f"C={C} {dtype}: rel err {e:.3e}"
    # the reference's own arithmetic (autocast) is no closer to fp32 than we are
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        amp = F.gelu(x + ffn1(x))
    assert e < 2.0 * _rel(amp.float(), ref) + 2e-4


def test_ffn1_against_the_reference_fixture():
    """tests/golden/pcblock.npz: inputs, parameters and `F.gelu(x + self.ffn1(x))` of the reference's own
    PCBlock4_Deep_nopool_res (fp32, CPU), plus the NumPy oracle on the same data."""
    import streamflow_b200 as sfb
    from oracle import streamflow_oracle as so
    from tests.helpers import load_golden, rel_err
    g = load_golden("pcblock.npz")
    C, H = g["w2"].shape
    ffn1 = nn.Sequential(nn.Conv2d(C, H, 1), nn.GELU(), nn.Conv2d(H, C, 1)).cuda().eval()
    with torch.no_grad():
        ffn1[0].weight.copy_(torch.from_numpy(g["w1"]).view(H, C, 1, 1)); ffn1[0].bias.copy_(torch.from_numpy(g["b1"]))
        ffn1[2].weight.copy_(torch.from_numpy(g["w2"]).view(C, H, 1, 1)); ffn1[2].bias.copy_(torch.from_numpy(g["b2"]))
        out = sfb.pcblock_ffn1(torch.from_numpy(g["x"]).cuda(), ffn1).cpu().numpy()
    assert rel_err(out, g["first"]) < 2e-3 and rel_err(out, so.pcblock_ffn1(g["x"], g["w1"], g["b1"], g["w2"], g["b2"])) < 2e-3


def test_ffn1_weight_update_and_errors():
    import streamflow_b200 as sfb
    ffn1 = _ffn1(128, seed=1)
    x = torch.randn(1, 128, 8, 8, device="cuda")
    with torch.no_grad():
        a = sfb.pcblock_ffn1(x, ffn1)
        ffn1[2].bias.add_(1.0)                       # in-place update bumps the version: the packed copy is refreshed
        b = sfb.pcblock_ffn1(x, ffn1)
        assert _rel(b, F.gelu(x + ffn1(x))) < 2e-3 and (a - b).abs().max() > 0.1
        with pytest.raises(sfb.StreamCorrError):
            sfb.pcblock_ffn1(x[:, :64], ffn1)
        with pytest.raises(sfb.StreamCorrError):
            sfb.pcblock_ffn1(torch.randn(1, 640, 8, 8, device="cuda"), _ffn1(640, seed=2))     # gru width: unsupported
        with pytest.raises(sfb.StreamCorrError):
            sfb.pcblock_ffn1(x.double(), ffn1)
    with torch.enable_grad(), pytest.raises(sfb.StreamCorrError):
        sfb.pcblock_ffn1(x.clone().requires_grad_(True), ffn1)                                  # inference only


def test_patched_motion_encoder_in_the_reference_model():
    """patch_motion_encoder() on the unmodified SKFlow_MF8 (oracle/_ref): the reference's own PCBlock under autocast vs the
    patched one on the tensor the real caller passes, then the final flows."""
    from oracle import ref_model as rm
    if not rm.available():
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py)")
    import streamflow_b200 as sfb
    from tests import model_harness as mh
    from tests.test_reference_model_gpu import _pair, _run
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, H, W = 4, 192, 320
    _, ours, _ = _pair(T, seed=7)
    frames = mh.synthetic_clip(T, H, W, seed=2)
    base = _run(ours, frames, 12)
    seen = {}
    blk = ours.update_block.encoder.convc1
    hook = blk.register_forward_pre_hook(lambda m, a: seen.setdefault("x", a[0].detach().clone()))
    _run(ours, frames, 1)
    hook.remove()
    x = seen["x"]
    assert x.shape[1] == 324 and x.dtype == torch.float32
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_blk = blk(x)
        names = sfb.patch_motion_encoder(ours)
        got_blk = blk(x)
    assert names == ["convc1", "convc2", "convf2", "conv"]
    e = _rel(got_blk.float(), ref_blk.float())
    assert got_blk.dtype == ref_blk.dtype and e < 5e-3, f"patched convc1 block: rel err {e:.3e}"
    n0 = sfb.lib().sf_launch_count()
    patched = _run(ours, frames, 12)
    assert sfb.lib().sf_launch_count() - n0 >= 12 * 4
    sfb.unpatch_motion_encoder(ours)
    for i in range(T - 1):
        epe = torch.sqrt(((patched[i] - base[i]) ** 2).sum(1)).mean().item()
        mag = torch.sqrt((base[i] ** 2).sum(1)).mean().item()
        print(f"pair {i}: |flow| {mag:.2f} px, mean EPE patched vs unpatched {epe:.5f} px")
        assert epe <= 0.01, f"pair {i}: mean EPE {epe:.4f} px"
    again = _run(ours, frames, 12)
    assert torch.equal(again[0], base[0])
