"""Pin the CPU oracle (NumPy restatement + torch CPU port) to the reference's outputs.

The golden ``.npz`` files were produced by importing the reference's ``core/corr.py`` and
``core/gma.py`` (tests/golden/make_golden.py).  Float tolerance: 2e-5 norm-wise for fp32
restatements (different summation order only); the coordinate-exact cases are tighter.
"""
import numpy as np
import pytest
import torch

from oracle import streamflow_oracle as so
from oracle import torch_port as tp
from tests.helpers import coord_sets, load_golden, rel_err, rs_normal

TOL = 2e-5
SETS = ["grid", "half", "jitter", "far", "neg", "border"]


@pytest.fixture(scope="module")
def small():
    return load_golden("corr_small.npz")


def test_pyramid_small_numpy(small):
    pyr = so.build_pyramid(small["f1"], small["f2"])
    for l in range(4):
        assert pyr[l].shape == small[f"level{l}"].shape
        assert rel_err(pyr[l], small[f"level{l}"]) < TOL
    # fp64 oracle agrees too (cross-check of the oracle itself)
    pyr64 = so.build_pyramid(small["f1"], small["f2"], dtype=np.float64)
    for l in range(4):
        assert rel_err(pyr64[l], small[f"level{l}"]) < TOL


@pytest.mark.parametrize("name", SETS)
def test_lookup_small_numpy(small, name):
    pyr = [small[f"level{l}"] for l in range(4)]
    got = so.lookup(pyr, small[f"coords_{name}"])
    ref = small[f"lookup_{name}"]
    assert got.shape == ref.shape == (1, 324, 17, 20) and got.dtype == np.float32
    assert rel_err(got, ref) < 1e-6
    # without the normalise/denormalise round trip the result moves by coordinate ulps only
    got2 = so.lookup(pyr, small[f"coords_{name}"], exact_roundtrip=False)
    assert rel_err(got2, ref) < 1e-4


def test_window_order_is_x_major(small):
    """Channel l*81 + i*9 + j samples x + (i-4), y + (j-4)."""
    pyr = [small[f"level{l}"] for l in range(4)]
    g = so.coords_grid(1, 17, 20)
    out = so.lookup(pyr, g)
    lvl0 = small["level0"].reshape(17, 20, 17, 20)
    y, x = 8, 9
    for (i, j) in [(0, 0), (8, 0), (0, 8), (5, 2), (4, 4)]:
        assert np.isclose(out[0, i * 9 + j, y, x], lvl0[y, x, y + j - 4, x + i - 4], rtol=1e-6)


def test_coords_sets_match_generator(small):
    mine = coord_sets(3, 1, 17, 20)
    for name in SETS:
        np.testing.assert_array_equal(mine[name], small[f"coords_{name}"])
    np.testing.assert_array_equal(so.coords_grid(1, 17, 20), small["coords_grid"])


def test_corr_batch_strided():
    g = load_golden("corr_batch.npz")
    f1 = np.transpose(g["f1_nhwc"], (0, 3, 1, 2))
    f2 = np.transpose(g["f2_nhwc"], (0, 3, 1, 2))
    pyr = so.build_pyramid(f1, f2)
    for l in range(4):
        assert rel_err(pyr[l], g[f"level{l}"]) < TOL
    assert rel_err(so.lookup(pyr, g["coords"]), g["lookup"]) < TOL
    # torch CPU port on the same strided views
    cp = tp.CpuCorrPyramid(torch.from_numpy(g["f1_nhwc"]).permute(0, 3, 1, 2),
                           torch.from_numpy(g["f2_nhwc"]).permute(0, 3, 1, 2))
    for l in range(4):
        assert rel_err(cp.levels[l].numpy(), g[f"level{l}"]) < TOL
    assert rel_err(cp(torch.from_numpy(g["coords"])).numpy(), g["lookup"]) < TOL


def test_corr_cfg1_known_answers():
    """BASELINE configs[0] (D=256, 46x62): sampled known answers + sums."""
    g = load_golden("corr_cfg1.npz")
    _, d, h, w = g["shape"]
    f1 = rs_normal(int(g["seed_f1"]), (1, d, h, w))
    f2 = rs_normal(int(g["seed_f2"]), (1, d, h, w))
    pyr = so.build_pyramid(f1, f2)
    shapes = [(2852, 1, 46, 62), (2852, 1, 23, 31), (2852, 1, 11, 15), (2852, 1, 5, 7)]
    for l in range(4):
        assert pyr[l].shape == shapes[l]
        flat = pyr[l].reshape(-1)
        assert rel_err(flat[g[f"level{l}_idx"]], g[f"level{l}_val"]) < TOL
        assert abs(np.abs(flat.astype(np.float64)).sum() / g[f"level{l}_abs"] - 1) < 1e-5
    port = tp.CpuCorrPyramid(torch.from_numpy(f1), torch.from_numpy(f2))
    for name, c in coord_sets(7, 1, h, w).items():
        flat = so.lookup(pyr, c).reshape(-1)
        assert rel_err(flat[g[f"lookup_{name}_idx"]], g[f"lookup_{name}_val"]) < TOL
        assert abs(np.abs(flat.astype(np.float64)).sum() / g[f"lookup_{name}_abs"] - 1) < 1e-5
        flat_t = port(torch.from_numpy(c)).numpy().reshape(-1)
        assert rel_err(flat_t[g[f"lookup_{name}_idx"]], g[f"lookup_{name}_val"]) < TOL


@pytest.mark.parametrize("name", ["gma_small.npz", "gma_proj.npz"])
def test_gma_numpy_and_port(name):
    g = load_golden(name)
    heads, dh = int(g["heads"]), int(g["dim_head"])
    attn = so.attention(g["inp"], g["w_qk"], heads=heads, dim_head=dh)
    assert attn.shape == g["attn"].shape
    assert rel_err(attn, g["attn"]) < TOL
    np.testing.assert_allclose(attn.sum(-1), 1.0, atol=1e-5)
    out = so.aggregate(attn, g["mf"], g["w_v"], float(g["gamma"]), g.get("w_proj"), heads=heads)
    assert rel_err(out, g["out"]) < TOL
    a_t = tp.cpu_attention(torch.from_numpy(g["inp"]), torch.from_numpy(g["w_qk"]), heads, dh)
    assert rel_err(a_t.numpy(), g["attn"]) < TOL
    wp = torch.from_numpy(g["w_proj"]) if "w_proj" in g else None
    o_t = tp.cpu_aggregate(a_t, torch.from_numpy(g["mf"]), torch.from_numpy(g["w_v"]), float(g["gamma"]), wp, heads)
    assert rel_err(o_t.numpy(), g["out"]) < TOL


def test_aggregate_gamma_zero_is_identity():
    """gamma is initialised to 0 in the reference (core/gma.py:84): output == fmap bit-exactly."""
    g = load_golden("gma_small.npz")
    out = so.aggregate(g["attn"], g["mf"], g["w_v"], 0.0)
    np.testing.assert_array_equal(out, g["mf"])


def test_pooling_is_linear():
    """pool(f1^T f2) == f1^T pool(f2): the identity the CUDA build kernel relies on."""
    f1, f2 = rs_normal(40, (1, 24, 10, 13)), rs_normal(41, (1, 24, 10, 13))
    pyr = so.build_pyramid(f1, f2, num_levels=3)
    f2p = so.avg_pool2x2(f2)
    vol1 = np.matmul(f1.reshape(1, 24, -1).transpose(0, 2, 1), f2p.reshape(1, 24, -1)) / np.sqrt(np.float32(24))
    assert rel_err(vol1.reshape(pyr[1].shape), pyr[1]) < TOL


def test_pcblock_entry_numpy():
    """oracle.pcblock_ffn1 against the value of `F.gelu(x + self.ffn1(x))` (core/update.py:31) computed by
    the reference's own PCBlock4_Deep_nopool_res (tests/golden/make_golden.py::case_pcblock)."""
    g = load_golden("pcblock.npz")
    out = so.pcblock_ffn1(g["x"], g["w1"], g["b1"], g["w2"], g["b2"])
    assert out.shape == g["first"].shape
    assert rel_err(out, g["first"]) < 2e-6          # the reference ran in fp32, the oracle in fp64
    # GELU restatement alone against torch's
    x = np.linspace(-6, 6, 97)
    assert np.allclose(so.gelu(x), torch.nn.functional.gelu(torch.from_numpy(x)).numpy(), atol=1e-12)
