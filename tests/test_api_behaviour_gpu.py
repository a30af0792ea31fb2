"""Library behaviour that only the real callers exercise (VERDICT r1 'weak' 13): autograd refusal,
inference_mode, two devices in one process, no device-wide side effects."""
import numpy as np
import pytest
import torch

from tests.helpers import rs_normal

pytestmark = pytest.mark.gpu


class _A:
    pass


def _mods(dev="cuda:0"):
    import streamflow_b200 as sfb
    att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
    agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
    with torch.no_grad():
        agg.gamma.fill_(0.5)
    return att, agg


def test_autograd_is_refused_not_silently_dropped():
    import streamflow_b200 as sfb
    att, agg = _mods()
    x = torch.randn(1, 128, 16, 24, device="cuda")
    f = torch.randn(1, 32, 16, 24, device="cuda")
    with torch.enable_grad():
        with pytest.raises(sfb.StreamCorrError, match="inference-only"):
            att(x)                                   # parameters require grad and autograd is on
        with pytest.raises(sfb.StreamCorrError, match="inference-only"):
            sfb.CorrBlock(f.clone().requires_grad_(), f)
    with torch.no_grad():
        h = att(x)
        blk = sfb.CorrBlock(f.clone().requires_grad_(), f)
        out = agg(h, x)
    assert not out.requires_grad and blk(sfb.coords_grid(1, 16, 24, device="cuda").contiguous()).shape == (1, 324, 16, 24)
    with torch.enable_grad():
        with pytest.raises(sfb.StreamCorrError, match="inference-only"):
            agg(h, x)
        for prm in list(att.parameters()) + list(agg.parameters()):
            prm.requires_grad_(False)
        assert agg(att(x), x).shape == x.shape       # frozen parameters: fine with autograd enabled


def test_inference_mode_both_conventions():
    import streamflow_b200 as sfb
    att, agg = _mods()
    attq = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128, return_qk=True).cuda()
    attq.load_state_dict(att.state_dict())
    x = torch.from_numpy(np.maximum(rs_normal(5, (2, 128, 12, 20)), 0)).cuda()
    mf = torch.from_numpy(rs_normal(6, (2, 128, 12, 20))).cuda()
    with torch.inference_mode():
        a = agg(att(x), mf)
        q, k = attq(x)                               # inference tensors: no version counter
        b = agg(q, k, mf)
        c = agg(q, k, mf)
    torch.cuda.synchronize()
    assert float((a - b).norm() / a.norm()) < 1e-3 and torch.equal(b, c)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """nn.DataParallel-style use (evaluate_mf.py:1207): the same process drives cuda:0 and cuda:1; per-device kernel
    attributes (dynamic shared memory opt-in) must be configured on each."""
    import streamflow_b200 as sfb
    outs = []
    f1, f2 = rs_normal(1, (1, 64, 24, 32)), rs_normal(2, (1, 64, 24, 32))
    inp, mf = np.maximum(rs_normal(3, (2, 128, 24, 32)), 0), rs_normal(4, (2, 128, 24, 32))
    att0, agg0 = _mods("cuda:0")
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        att, agg = _mods(dev)
        att.load_state_dict(att0.state_dict())
        agg.load_state_dict(agg0.state_dict())
        with torch.no_grad():                        # current device stays cuda:0 throughout: tensors choose
            blk = sfb.CorrBlock(torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev))
            feats = blk(sfb.coords_grid(1, 24, 32, device=dev).contiguous() + 0.3)
            out = agg(att(torch.from_numpy(inp).to(dev)), torch.from_numpy(mf).to(dev))
        torch.cuda.synchronize(dev)
        outs.append((feats.cpu(), out.cpu()))
    for feats, out in outs[1:]:
        assert torch.equal(feats, outs[0][0]) and torch.equal(out, outs[0][1])     # deterministic across devices


def test_no_device_wide_limits_are_changed():
    """Loading and using the library must not alter cudaLimitMaxL2FetchGranularity (it would also apply to the
    cuDNN / cuBLAS kernels of the unchanged update block)."""
    import ctypes
    import streamflow_b200 as sfb
    rt = ctypes.CDLL("libcudart.so.12")
    before = ctypes.c_size_t()
    assert rt.cudaDeviceGetLimit(ctypes.byref(before), 5) == 0       # cudaLimitMaxL2FetchGranularity = 0x05
    f = torch.randn(1, 32, 16, 24, device="cuda")
    with torch.no_grad():
        sfb.CorrBlock(f, f)
    torch.cuda.synchronize()
    after = ctypes.c_size_t()
    assert rt.cudaDeviceGetLimit(ctypes.byref(after), 5) == 0
    assert before.value == after.value
