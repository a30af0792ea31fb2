"""GPU parity: CUDA CorrBlock (through the C ABI) vs the CPU oracle and the reference's golden vectors.

Tolerances (norm-wise relative error ||a-b||/||b||, the norm BASELINE.json's north_star refers to):
  lookup on a given pyramid ........ 1e-5  (fp32 bilinear; only the coordinate round trip differs)
  build, precision 'fp32' .......... 1e-5  (fp32 FFMA, different summation order)
  build, precision 'f16x2' ......... 2e-5  (split fp16 operands, fp32 accumulate)
  build, precision 'f16' ........... 1e-3  (north_star bound; operands rounded to 11-bit mantissa)
"""
import numpy as np
import pytest
import torch

from oracle import streamflow_oracle as so
from tests.helpers import coord_sets, load_golden, max_rel_err, rel_err, rs_normal

pytestmark = pytest.mark.gpu

TOL_BUILD = {"fp32": 1e-5, "f16x2": 2e-5, "f16": 1e-3, "auto": 2e-5}
SETS = ["grid", "half", "jitter", "far", "neg", "border"]


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def small():
    return load_golden("corr_small.npz")


@pytest.mark.parametrize("variant", ["reg", "ws"])
@pytest.mark.parametrize("name", SETS)
def test_lookup_on_reference_pyramid(small, name, variant, monkeypatch):
    """Both lookup kernels (register-staged for short launches, warp-specialised cp.async for long ones; the library picks by launch
    size, STREAMCORR_LOOKUP forces one) against the reference's grid_sample lookup on the reference's own pyramid."""
    from streamflow_b200 import CorrBlock
    monkeypatch.setenv("STREAMCORR_LOOKUP", variant)
    blk = CorrBlock.from_dense_pyramid([cuda(small[f"level{l}"]) for l in range(4)])
    out = blk(cuda(small[f"coords_{name}"]))
    assert out.shape == (1, 324, 17, 20) and out.dtype == torch.float32 and out.is_contiguous()
    err = rel_err(out.cpu().numpy(), small[f"lookup_{name}"])
    assert err < 1e-5, f"lookup[{name}] rel err {err:.3e}"


@pytest.mark.parametrize("prec", ["fp32", "f16x2", "f16", "auto"])
def test_build_small_vs_reference(small, prec):
    from streamflow_b200 import CorrBlock
    blk = CorrBlock(cuda(small["f1"]), cuda(small["f2"]), precision=prec)
    for l in range(4):
        got = blk.corr_pyramid[l].cpu().numpy()
        assert got.shape == small[f"level{l}"].shape
        err = rel_err(got, small[f"level{l}"])
        assert err < TOL_BUILD[prec], f"{prec} level {l}: rel err {err:.3e}"
    # pad cells of the 4x4-tiled storage are zero (the lookup relies on it)
    for buf, (hl, wl, th, tw) in zip(blk._levels, blk._dims):
        img = buf.view(-1, th, tw, 4, 4).permute(0, 1, 3, 2, 4).reshape(-1, th * 4, tw * 4)
        if th * 4 != hl:
            assert float(img[:, hl:, :].abs().max()) == 0.0
        if tw * 4 != wl:
            assert float(img[:, :, wl:].abs().max()) == 0.0
    out = blk(cuda(small["coords_jitter"]))
    assert rel_err(out.cpu().numpy(), small["lookup_jitter"]) < TOL_BUILD[prec]


@pytest.mark.parametrize("prec", ["fp32", "f16x2", "f16", "auto"])
def test_build_batch_strided(prec):
    """B=2, channels-last strided inputs exactly as the model passes them (streamflow.py:107,110)."""
    from streamflow_b200 import CorrBlock
    g = load_golden("corr_batch.npz")
    f1 = cuda(g["f1_nhwc"]).permute(0, 3, 1, 2)
    f2 = cuda(g["f2_nhwc"]).permute(0, 3, 1, 2)
    assert not f1.is_contiguous()
    blk = CorrBlock(f1, f2, precision=prec)
    for l in range(4):
        err = rel_err(blk.corr_pyramid[l].cpu().numpy(), g[f"level{l}"])
        assert err < TOL_BUILD[prec], f"{prec} level {l}: rel err {err:.3e}"
    err = rel_err(blk(cuda(g["coords"])).cpu().numpy(), g["lookup"])
    assert err < TOL_BUILD[prec], f"{prec} lookup: {err:.3e}"


@pytest.mark.parametrize("prec", ["fp32", "f16x2", "f16", "auto"])
def test_cfg1_known_answers(prec):
    """BASELINE configs[0]: D=256, 46x62 (ragged: w=62 -> pitch 64, odd pooled sizes)."""
    from streamflow_b200 import CorrBlock
    g = load_golden("corr_cfg1.npz")
    _, d, h, w = [int(v) for v in g["shape"]]
    f1, f2 = rs_normal(int(g["seed_f1"]), (1, d, h, w)), rs_normal(int(g["seed_f2"]), (1, d, h, w))
    blk = CorrBlock(cuda(f1), cuda(f2), precision=prec)
    for l in range(4):
        flat = blk.corr_pyramid[l].contiguous().cpu().numpy().reshape(-1)
        err = rel_err(flat[g[f"level{l}_idx"]], g[f"level{l}_val"])
        assert err < TOL_BUILD[prec], f"{prec} level {l}: {err:.3e}"
        assert abs(np.abs(flat.astype(np.float64)).sum() / float(g[f"level{l}_abs"]) - 1) < TOL_BUILD[prec]
    for name, c in coord_sets(7, 1, h, w).items():
        flat = blk(cuda(c)).cpu().numpy().reshape(-1)
        err = rel_err(flat[g[f"lookup_{name}_idx"]], g[f"lookup_{name}_val"])
        assert err < TOL_BUILD[prec], f"{prec} lookup[{name}]: {err:.3e}"


def test_fp16_representable_inputs_are_exact_products():
    """With fp16-representable feature maps (the model's mixed-precision path) the f16 mode has no operand
    rounding at all: it must agree with the fp32 oracle to accumulation-order noise."""
    from streamflow_b200 import CorrBlock
    f1 = rs_normal(50, (1, 128, 24, 32)).astype(np.float16).astype(np.float32)
    f2 = rs_normal(51, (1, 128, 24, 32)).astype(np.float16).astype(np.float32)
    pyr = so.build_pyramid(f1, f2)
    blk = CorrBlock(cuda(f1), cuda(f2), precision="f16")
    err0 = rel_err(blk.corr_pyramid[0].cpu().numpy(), pyr[0])
    assert err0 < 2e-6, f"level 0: {err0:.3e}"
    # pooled operands are rounded once more (avg of 4 fp16 values needs 2 extra bits)
    for l in range(1, 4):
        assert rel_err(blk.corr_pyramid[l].cpu().numpy(), pyr[l]) < 5e-4


def test_auto_precision_picks_the_path_on_the_device():
    """Default precision 'auto' (SF_PREC_AUTO): the absmax pass finds out on the device whether the feature maps are
    fp16-representable.  Exact inputs -> single-product level 0 + hi*hi + hi*lo pooled levels; arbitrary fp32 inputs
    -> the three-product path.  Both are fp32-faithful: ALL levels agree with the fp32 oracle to ~1e-6 / 2e-5."""
    from streamflow_b200 import CorrBlock
    f1 = rs_normal(60, (1, 128, 24, 32))
    f2 = rs_normal(61, (1, 128, 24, 32))
    e1, e2 = f1.astype(np.float16).astype(np.float32), f2.astype(np.float16).astype(np.float32)
    for a, b, tol, what in ((e1, e2, 5e-6, "fp16-exact inputs"), (f1, f2, 2e-5, "arbitrary fp32 inputs"),
                            (e1, f2, 2e-5, "one inexact operand")):
        pyr = so.build_pyramid(a, b)
        blk = CorrBlock(cuda(a), cuda(b))                    # default = auto
        assert blk.precision == "auto"
        for l in range(4):
            err = rel_err(blk.corr_pyramid[l].cpu().numpy(), pyr[l])
            assert err < tol, f"{what}, level {l}: {err:.3e}"
    # D not a multiple of 64: auto falls back to the three-product mode (still fp32-faithful)
    g1, g2 = rs_normal(62, (1, 40, 16, 20)), rs_normal(63, (1, 40, 16, 20))
    blk = CorrBlock(cuda(g1), cuda(g2))
    pyr = so.build_pyramid(g1, g2)
    for l in range(4):
        assert rel_err(blk.corr_pyramid[l].cpu().numpy(), pyr[l]) < 2e-5
    # strided channels-last views of a clip (what the model hands over), batched group build
    from streamflow_b200 import CorrGroup
    fm = rs_normal(64, (1, 3, 24, 32, 64)).astype(np.float16).astype(np.float32)
    fmaps = cuda(fm).permute(0, 1, 4, 2, 3)
    grp = CorrGroup.from_fmaps(fmaps)
    nchw = np.transpose(fm, (0, 1, 4, 2, 3))
    for i in range(2):
        pyr = so.build_pyramid(nchw[:, i], nchw[:, i + 1])
        for l in range(4):
            assert rel_err(grp.blocks[i].corr_pyramid[l].cpu().numpy(), pyr[l]) < 5e-6


def test_operand_scaling_extreme_ranges():
    """Per-tensor power-of-two scaling: inputs far outside the fp16 range still give 1e-3 parity."""
    from streamflow_b200 import CorrBlock
    f1 = rs_normal(52, (1, 64, 16, 16)) * np.float32(3e4)
    f2 = rs_normal(53, (1, 64, 16, 16)) * np.float32(2e-6)
    pyr = so.build_pyramid(f1, f2)
    for prec in ["f16", "f16x2", "auto"]:
        blk = CorrBlock(cuda(f1), cuda(f2), precision=prec)
        for l in range(4):
            err = rel_err(blk.corr_pyramid[l].cpu().numpy(), pyr[l])
            assert err < TOL_BUILD[prec], f"{prec} level {l}: {err:.3e}"


def test_sintel_size_properties():
    """BASELINE configs[1] size (55x128, D=256): properties that need no CPU volume.

    (1) symmetry  corr(f1,f2)[n,m] == corr(f2,f1)[m,n]   (the reference's own __main__ check, corr.py:56-68)
    (2) lookup at integer coordinates returns volume entries verbatim (zero outside)
    (3) each pooled level equals avg_pool2d of the level above within the f16-mode bound
    """
    from streamflow_b200 import CorrBlock, coords_grid
    torch.manual_seed(0)
    h, w, d = 55, 128, 256
    fm = torch.randn(1, 2, h, w, d, device="cuda").half().float().permute(0, 1, 4, 2, 3)  # channels-last views
    f1, f2 = fm[:, 0], fm[:, 1]
    a = CorrBlock(f1, f2)
    b = CorrBlock(f2, f1)
    v_ab = a.corr_pyramid[0][:, 0].reshape(h * w, h * w)
    v_ba = b.corr_pyramid[0][:, 0].reshape(h * w, h * w)
    assert float((v_ab - v_ba.t()).abs().max()) <= 1e-4 * float(v_ab.abs().max())
    # (2)
    g = coords_grid(1, h, w, device="cuda").contiguous()
    out = a(g)
    vol = v_ab.view(h, w, h, w)
    for (i, j, y, x) in [(4, 4, 10, 20), (0, 0, 30, 100), (8, 8, 54, 127), (8, 0, 0, 0), (2, 7, 27, 64)]:
        yy, xx = y + j - 4, x + i - 4
        ref = float(vol[y, x, yy, xx]) if (0 <= yy < h and 0 <= xx < w) else 0.0
        assert abs(float(out[0, i * 9 + j, y, x]) - ref) <= 1e-6 * max(1.0, abs(ref))
    # (3)
    for l in range(3):
        pooled = torch.nn.functional.avg_pool2d(a.corr_pyramid[l], 2, stride=2)
        nxt = a.corr_pyramid[l + 1]
        err = float((pooled - nxt).norm() / pooled.norm())
        assert err < 1e-3, f"level {l + 1} vs pooled level {l}: {err:.3e}"
    # level sums: checksum of checksums (linearity of the GEMM in f2)
    s_direct = float(a.corr_pyramid[0].double().sum())
    s_lin = float((f1.double().sum(dim=(2, 3)) * f2.double().sum(dim=(2, 3))).sum() / 16.0)
    assert abs(s_direct - s_lin) <= 2e-3 * max(1.0, abs(s_lin)) + 1e-3 * float(a.corr_pyramid[0].abs().double().sum()) * 1e-3


def test_error_behaviour():
    from streamflow_b200 import CorrBlock, StreamCorrError
    x = torch.zeros(1, 8, 16, 16, device="cuda")
    with pytest.raises(StreamCorrError):
        CorrBlock(x, x, num_levels=3)
    with pytest.raises(StreamCorrError):
        CorrBlock(x, x, radius=3)
    with pytest.raises(StreamCorrError):
        CorrBlock(x, torch.zeros(1, 8, 16, 17, device="cuda"))
    with pytest.raises(StreamCorrError):
        CorrBlock(torch.zeros(1, 8, 8, 8, device="cuda"), torch.zeros(1, 8, 8, 8, device="cuda"))  # 8>>3 = 1
    blk = CorrBlock(x, x, precision="fp32")
    with pytest.raises(StreamCorrError):
        blk(torch.zeros(1, 2, 16, 15, device="cuda"))


def test_lookup_kernels_agree_bit_for_bit_under_stress(monkeypatch):
    """The warp-specialised cp.async kernel (loader and interpolation warps coupled only by mbarriers; compute-sanitizer's
    racecheck does not model cp.async.mbarrier.arrive and flags the staging buffers) against the register-staged kernel
    (plain __syncthreads): the same arithmetic on the same tiles, so any missed hand-over would show as a bit difference.
    3 Sintel-size pairs (4 work items per CTA, every buffer reused), 40 coordinate sets incl. far out-of-bounds ones."""
    from streamflow_b200 import CorrBlock, CorrGroup
    torch.manual_seed(11)
    h, w = 55, 128
    fm = torch.randn(1, 4, h, w, 64, device="cuda").half().float().permute(0, 1, 4, 2, 3)
    group = CorrGroup([CorrBlock(fm[:, i], fm[:, i + 1], radius=4) for i in range(3)])
    ys, xs = torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij")
    grid = torch.stack((xs, ys), 0).float()[None]
    for it in range(40):
        scale = [1.0, 6.0, 40.0, 300.0][it % 4]
        coords = [(grid + scale * torch.randn(1, 2, h, w, device="cuda")).contiguous() for _ in range(3)]
        monkeypatch.setenv("STREAMCORR_LOOKUP", "ws")
        a = [t.clone() for t in group(coords)]
        monkeypatch.setenv("STREAMCORR_LOOKUP", "reg")
        b = group(coords)
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"iteration {it}: kernels differ in {(x != y).sum().item()} values"
