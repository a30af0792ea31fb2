"""SURVEY 8(f) "next" rows: convex upsampling kernel (GPU), .flo wire format and InputPadder (CPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from streamflow_b200 import flowio
from tests.helpers import rs_normal


def ref_upsample(flow, mask, ratio=8):
    """core/models/streamflow.py:82-93 restated."""
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, ratio, ratio, h, w).float(), dim=2)
    up = F.unfold(ratio * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, ratio * h, ratio * w)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((1, 55, 128), torch.float32), ((3, 55, 128), torch.float16),
                                         ((2, 23, 37), torch.float32), ((1, 5, 3), torch.bfloat16)])
def test_upsample_flow_matches_reference(shape, dtype):
    from streamflow_b200 import upsample_flow
    n, h, w = shape
    flow = torch.from_numpy(rs_normal(80, (n, 2, h, w)) * 4).cuda()
    mask = (torch.from_numpy(rs_normal(81, (n, 576, h, w))) * 2).cuda().to(dtype)
    got = upsample_flow(flow, mask)
    want = ref_upsample(flow, mask)
    assert got.shape == (n, 2, 8 * h, 8 * w) and got.dtype == torch.float32
    err = float((got - want).norm() / want.norm())
    assert err < 1e-5, f"rel err {err:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float16, 2e-3)])
def test_upsample_flow_matches_reference_fixture(dtype, tol):
    """tests/golden/upsample.npz was produced by the reference METHOD SKFlow_MF8.upsample_flow itself
    (core/models/streamflow.py:82-93, loaded through oracle/ref_model.py by tests/golden/make_golden.py)."""
    from streamflow_b200 import upsample_flow
    from tests.helpers import load_golden
    g = load_golden("upsample.npz")
    flow, mask = torch.from_numpy(g["flow"]).cuda(), torch.from_numpy(g["mask"]).cuda().to(dtype)
    got = upsample_flow(flow, mask).cpu()
    want = torch.from_numpy(g["out"])
    assert got.shape == want.shape
    err = float((got - want).norm() / want.norm())
    assert err < tol, f"rel err {err:.3e}"


def test_upsample_fixture_agrees_with_the_restatement():
    """CPU: the test-side restatement used for the larger GPU shapes equals the reference-generated fixture."""
    from tests.helpers import load_golden
    g = load_golden("upsample.npz")
    want = ref_upsample(torch.from_numpy(g["flow"]), torch.from_numpy(g["mask"]))
    np.testing.assert_allclose(want.numpy(), g["out"], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_patch_upsample_replaces_method():
    from streamflow_b200 import patch_upsample

    class Dummy:
        def upsample_flow(self, flow, mask, ratio=8):
            raise AssertionError("should have been patched")

    patch_upsample(Dummy)
    flow = torch.zeros(1, 2, 4, 4, device="cuda")
    mask = torch.zeros(1, 576, 4, 4, device="cuda")
    assert Dummy().upsample_flow(flow, mask).shape == (1, 2, 32, 32)


def test_flo_roundtrip_and_layout(tmp_path):
    flow = rs_normal(90, (2, 7, 11))
    p = tmp_path / "a.flo"
    flowio.write_flo(p, torch.from_numpy(flow))
    raw = np.fromfile(p, np.uint8)
    assert raw.size == 12 + 7 * 11 * 2 * 4
    assert np.frombuffer(raw[:4].tobytes(), np.float32)[0] == np.float32(202021.25)
    assert tuple(np.frombuffer(raw[4:12].tobytes(), np.int32)) == (11, 7)          # width, then height
    body = np.frombuffer(raw[12:].tobytes(), np.float32).reshape(7, 11, 2)          # interleaved u, v
    np.testing.assert_array_equal(body[..., 0], flow[0])
    np.testing.assert_array_equal(body[..., 1], flow[1])
    back = flowio.read_flo(p)
    np.testing.assert_array_equal(back, np.transpose(flow, (1, 2, 0)))
    (tmp_path / "bad.flo").write_bytes(b"\\x00" * 32)
    with pytest.raises(ValueError):
        flowio.read_flo(tmp_path / "bad.flo")


def test_input_padder_sintel_and_kitti():
    x = torch.arange(436 * 1024, dtype=torch.float32).view(1, 1, 436, 1024)
    pad = flowio.InputPadder(x.shape)                       # sintel: 436 -> 440, two rows top and bottom
    y, = pad.pad(x)
    assert y.shape[-2:] == (440, 1024) and pad._pad == [0, 0, 2, 2]
    assert torch.equal(y[..., 0, :], x[..., 0, :]) and torch.equal(y[..., -1, :], x[..., -1, :])   # replicate
    assert torch.equal(pad.unpad(y), x)
    k = torch.zeros(1, 3, 375, 1242)
    padk = flowio.InputPadder(k.shape, mode="kitti")        # bottom-only vertical padding
    assert padk._pad == [3, 3, 0, 1] and padk.pad(k)[0].shape[-2:] == (376, 1248)
    assert padk.unpad(padk.pad(k)[0]).shape == k.shape
    assert flowio.InputPadder((376, 1248))._pad == [0, 0, 0, 0]


def test_kitti_flow_png_layout_and_reference_semantics(tmp_path):
    """KITTI 16-bit flow PNG (core/utils/frame_utils.py:118-123, 137-141): R = 64u + 2^15, G = 64v + 2^15, B = valid;
    byte layout checked directly, pixel equality with the reference's cv2 writer / reader when cv2 is present."""
    import struct
    import zlib
    import numpy as np
    from streamflow_b200.flowio import read_flow_kitti, write_flow_kitti
    rs = np.random.RandomState(3)
    uv = (rs.standard_normal((19, 23, 2)) * 30).astype(np.float32)
    valid = rs.uniform(size=(19, 23)) > 0.25
    p = str(tmp_path / "flow.png")
    write_flow_kitti(p, torch.from_numpy(uv).permute(2, 0, 1), valid)          # [2, H, W] tensor accepted
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and raw[12:16] == b"IHDR"
    w, h, depth, ctype = struct.unpack(">IIBB", raw[16:26])
    assert (w, h, depth, ctype) == (23, 19, 16, 2)
    n = struct.unpack(">I", raw[33:37])[0]
    assert raw[37:41] == b"IDAT"
    lines = np.frombuffer(zlib.decompress(raw[41:41 + n]), np.uint8).reshape(19, 1 + 23 * 6)
    assert (lines[:, 0] == 0).all()
    px = lines[:, 1:].copy().view(">u2").reshape(19, 23, 3)
    want = (64.0 * uv + 2 ** 15).astype(np.uint16)                              # the reference's quantisation
    assert np.array_equal(px[..., :2], want) and np.array_equal(px[..., 2] != 0, valid)
    flow, v = read_flow_kitti(p)
    assert flow.dtype == np.float32 and np.abs(flow - uv).max() <= 1 / 64 and np.array_equal(v != 0, valid)
    with pytest.raises(ValueError):
        write_flow_kitti(p, np.full((4, 4, 2), 600.0, np.float32))              # outside the 16-bit range
    cv2 = pytest.importorskip("cv2")
    ref = cv2.imread(p, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float32)   # readFlowKITTI
    assert np.array_equal((ref[:, :, :2] - 2 ** 15) / 64.0, flow) and np.array_equal(ref[:, :, 2], v)
    p2 = str(tmp_path / "ref.png")                                              # writeFlowKITTI -> our reader
    arr = np.concatenate([64.0 * uv + 2 ** 15, np.ones((19, 23, 1))], axis=-1).astype(np.uint16)
    cv2.imwrite(p2, arr[..., ::-1])
    flow2, v2 = read_flow_kitti(p2)                                             # cv2 uses PNG filters: exercises them
    assert np.array_equal(flow2, flow) and (v2 == 1).all()
