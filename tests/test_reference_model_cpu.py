"""The UNMODIFIED reference model (oracle/_ref: core/models/streamflow.py + core/update.py + core/encoders) loads
around both sets of hot-path operators; CPU-only checks (construction, parameter names, a tiny reference forward).
The GPU half is tests/test_reference_model_gpu.py."""
import sys
import warnings

import pytest
import torch

from oracle import ref_model as rm

pytestmark = pytest.mark.skipif(not rm.available(), reason="oracle/_ref not built (python oracle/make_ref.py)")


@pytest.fixture(scope="module")
def modules():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return rm.load_model_module("reference"), rm.load_model_module("b200")


def test_install_binds_the_reference_imports(modules):
    import streamflow_b200 as sfb
    ref, ours = modules
    # `from corr import CorrBlock` (streamflow.py:8), `from gma import Attention` (:10), update.py:4 `Aggregate`
    assert ours.CorrBlock is sfb.CorrBlock and ours.Attention is sfb.Attention
    assert ours._l1_modules["update"].Aggregate is sfb.Aggregate
    assert ref.CorrBlock.__module__ == "corr" and ref.CorrBlock is not sfb.CorrBlock
    assert ref._l1_modules["update"].Aggregate.__module__ == "gma"
    # the two variants are distinct module objects and leave sys.modules clean
    assert ours is not ref and ours._l1_modules["update"] is not ref._l1_modules["update"]
    assert not any(k in sys.modules for k in ("corr", "gma", "update"))


def test_shipped_model_builds_on_both_and_shares_checkpoints(modules):
    import streamflow_b200 as sfb
    ref_mod, our_mod = modules
    torch.manual_seed(0)
    ref = rm.randomise(rm.build_model(ref_mod), seed=1)
    ours = rm.build_model(our_mod)
    assert type(ours.update_block.aggregator) is sfb.Aggregate and type(ours.att) is sfb.Attention
    assert type(ours.update_block).__name__ == "SKUpdateBlock_TAM_v3" and type(ours.fnet).__name__ == "Twins_CSC"
    assert ours.load_state_dict(ref.state_dict(), strict=True).missing_keys == []
    assert ref.load_state_dict(ours.state_dict(), strict=True).missing_keys == []
    keys = set(ref.state_dict())
    assert {"att.to_qk.weight", "update_block.aggregator.to_v.weight", "update_block.aggregator.gamma"} <= keys
    # StreamFlow reports ~14.2 M parameters; the timm shim restates twins_svt_large stages 0-1 with timm's shapes
    n = sum(p.numel() for p in ref.parameters())
    assert 14.0e6 < n < 14.5e6, n
    assert float(ref.update_block.aggregator.gamma) != 0.0


def test_reference_forward_runs_on_cpu(modules):
    ref_mod, _ = modules
    torch.manual_seed(0)
    model = rm.randomise(rm.build_model(ref_mod), seed=1).eval()
    g = torch.Generator().manual_seed(0)
    frames = [torch.randint(0, 256, (1, 3, 128, 192), generator=g).float() for _ in range(4)]
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        flows = model(frames, iters=2, test_mode=True)
    assert len(flows) == 3 and all(f.shape == (1, 2, 128, 192) and torch.isfinite(f).all() for f in flows)


def test_input_padder_matches_the_reference_padder():
    """streamflow_b200.flowio.InputPadder (own implementation) pads / unpads exactly like core/utils/utils.py:7-31."""
    import importlib.util
    import os
    from oracle.make_ref import ref_core_dir
    from streamflow_b200.flowio import InputPadder as Ours
    core = ref_core_dir()
    sys.path.insert(0, core)
    try:
        spec = importlib.util.spec_from_file_location("_ref_utils", os.path.join(core, "utils", "utils.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(core)
    for dims in [(436, 1024), (376, 1248), (1080, 1920), (37, 53), (440, 1024), (1, 3, 375, 1242)]:
        for mode in ("sintel", "kitti"):
            a, b = Ours(dims, mode), mod.InputPadder(dims, mode)
            assert a._pad == b._pad, (dims, mode)
            x = torch.arange(3 * dims[-2] * dims[-1], dtype=torch.float32).view(1, 3, dims[-2], dims[-1])
            pa, pb = a.pad(x)[0], b.pad(x)[0]
            assert torch.equal(pa, pb) and torch.equal(a.unpad(pa), x)
