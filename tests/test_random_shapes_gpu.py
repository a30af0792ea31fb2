"""Randomised small shapes through the C ABI against the NumPy oracle: ragged sizes, batches, strides, dtypes.
(Deterministic: shapes are drawn from a seeded generator so failures reproduce.)"""
import numpy as np
import pytest
import torch

from oracle import streamflow_oracle as so
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "f16x2": 2e-5, "f16": 1e-3}


def cases(n, seed=123):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        B = int(rng.randint(1, 4))
        D = int(rng.choice([8, 24, 40, 64, 72, 128]))
        h = int(rng.randint(16, 41))
        w = int(rng.randint(16, 49))
        layout = str(rng.choice(["nchw", "nhwc", "sliced"]))
        prec = str(rng.choice(["f16", "f16x2", "fp32"]))
        out.append((B, D, h, w, layout, prec, int(rng.randint(0, 1 << 30))))
    return out


@pytest.mark.parametrize("B,D,h,w,layout,prec,seed", cases(14))
def test_build_and_lookup_random_shapes(B, D, h, w, layout, prec, seed):
    from streamflow_b200 import CorrBlock, CorrGroup
    rs = np.random.RandomState(seed)
    f = rs.standard_normal((B, 2, D, h, w)).astype(np.float32)
    t = torch.from_numpy(f).cuda()
    if layout == "nhwc":          # channels-last storage viewed as NCHW, like the model's encoder output
        t = t.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    elif layout == "sliced":      # non-dense view: every other channel of a 2D-channel tensor
        big = torch.zeros(B, 2, 2 * D, h, w, device="cuda")
        big[:, :, ::2] = t
        t = big[:, :, ::2]
    f1, f2 = t[:, 0], t[:, 1]
    blk = CorrBlock(f1, f2, precision=prec)
    pyr = so.build_pyramid(f[:, 0], f[:, 1])
    for l in range(4):
        err = rel_err(blk.corr_pyramid[l].cpu().numpy(), pyr[l])
        assert err < TOL[prec], f"level {l}: {err:.3e}"
    coords = so.coords_grid(B, h, w) + rs.uniform(-1.5 * max(h, w), 1.5 * max(h, w), (B, 2, h, w)).astype(np.float32) * \
        (rs.uniform(size=(B, 1, h, w)) < 0.3) + 2.0 * rs.standard_normal((B, 2, h, w)).astype(np.float32)
    coords = coords.astype(np.float32)
    want = so.lookup(pyr, coords)
    got = blk(torch.from_numpy(coords).cuda()).cpu().numpy()
    assert rel_err(got, want) < TOL[prec]
    # non-contiguous / fp64 coordinates are accepted like the reference accepts any float tensor
    c64 = torch.from_numpy(coords).cuda().double()
    assert rel_err(blk(c64).cpu().numpy(), want) < TOL[prec]
    grp = CorrGroup([blk, blk])([torch.from_numpy(coords).cuda()] * 2).view(B, 2, 324, h, w)
    assert rel_err(grp[:, 0].cpu().numpy(), want) < TOL[prec] and torch.equal(grp[:, 0], grp[:, 1])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_low_precision_feature_maps_are_upcast(dtype):
    from streamflow_b200 import CorrBlock
    rs = np.random.RandomState(5)
    f = torch.from_numpy(rs.standard_normal((1, 2, 64, 20, 28)).astype(np.float32)).cuda().to(dtype)
    blk = CorrBlock(f[:, 0], f[:, 1])
    pyr = so.build_pyramid(f[:, 0].float().cpu().numpy(), f[:, 1].float().cpu().numpy())
    assert rel_err(blk.corr_pyramid[0].cpu().numpy(), pyr[0]) < 1e-3


# The aggregate splits the query rows of all maps into 16-row units over the SMs; the extra cases pin its partition
# edges: more maps than SMs with one unit each (160 x 16), a ragged last unit (N = 20), N not a multiple of 4
# (scalar epilogue, N = 15), several CTAs per map with uneven runs (N = 156, 2600).
@pytest.mark.parametrize("P,h,w,seed", [(1, 9, 13, 1), (2, 16, 16, 2), (3, 11, 24, 3), (1, 40, 33, 4),
                                        (160, 4, 4, 5), (5, 4, 5, 6), (1, 3, 5, 7), (7, 12, 13, 8), (2, 50, 52, 9)])
def test_gma_random_shapes(P, h, w, seed):
    from streamflow_b200 import Aggregate, Attention

    class A:
        pass
    rs = np.random.RandomState(seed)
    inp = np.maximum(rs.standard_normal((P, 128, h, w)), 0).astype(np.float32)
    mf = rs.standard_normal((P, 128, h, w)).astype(np.float32)
    w_qk = (rs.standard_normal((256, 128)) * 0.2).astype(np.float32)
    w_v = (rs.standard_normal((128, 128)) * 0.1).astype(np.float32)
    att = Attention(args=A(), dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg = Aggregate(args=A(), dim=128, heads=1, dim_head=128).cuda()
    with torch.no_grad():
        att.to_qk.weight.copy_(torch.from_numpy(w_qk).view(256, 128, 1, 1))
        agg.to_v.weight.copy_(torch.from_numpy(w_v).view(128, 128, 1, 1))
        agg.gamma.fill_(-0.6)
    hd = att(torch.from_numpy(inp).cuda())
    attn = so.attention(inp, w_qk)
    assert rel_err(hd.dense().cpu().numpy(), attn) < 1e-3
    want = so.aggregate(attn, mf, w_v, -0.6)
    for _ in range(2):      # second call exercises the re-zeroed accumulator
        got = agg(hd, torch.from_numpy(mf).cuda()).cpu().numpy()
        assert rel_err(got - mf, want - mf) < 1e-3
    # channels-last motion features (a non-contiguous view) are accepted
    mf_cl = torch.from_numpy(mf).cuda().contiguous(memory_format=torch.channels_last)
    assert rel_err(agg(hd, mf_cl).cpu().numpy() - mf, want - mf) < 1e-3
