"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/streamcorr.h
declares, answers its pure-host queries, and refuses to compute without an sm_100 device (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from streamflow_b200.build import build_library
    build_library()
    import streamflow_b200
    return streamflow_b200.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "streamcorr.h")).read()
    return sorted(set(re.findall(r"SF_API\s+[\w\s\*]+?\b(sf_\w+)\s*\(", text)))


def test_exports_every_header_symbol(L):
    from streamflow_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 14
    for name in syms:
        assert hasattr(L, name), f"libstreamcorr.so does not export {name}"
    assert sorted(_lib.EXPORTS) == syms, "streamflow_b200/_lib.py EXPORTS out of sync with the header"


def test_version_and_level_geometry(L):
    from streamflow_b200 import _lib
    assert L.sf_version() == 210
    # Sintel 55x128 and BASELINE configs[0] 46x62 (floor-mode pooling, pitch rounded up to 4 floats)
    assert [_lib.level_dims(55, 128, l) for l in range(4)] == [(55, 128, 14, 32), (27, 64, 7, 16), (13, 32, 4, 8),
                                                               (6, 16, 2, 4)]
    assert [_lib.level_dims(46, 62, l) for l in range(4)] == [(46, 62, 12, 16), (23, 31, 6, 8), (11, 15, 3, 4),
                                                              (5, 7, 2, 2)]
    assert L.sf_gma_npad(7040) == 7040 and L.sf_gma_npad(2852) == 2880
    assert L.sf_gma_e_elems(3, 7040) == 3 * 7040 * 7040
    assert L.sf_gma_e_elems(1, 2852) == 2944 * 2880


def test_workspace_queries(L):
    from streamflow_b200 import _lib
    n = 55 * 128
    w16 = L.sf_corr_workspace_bytes(1, 256, 55, 128, _lib.PREC_F16)
    w32 = L.sf_corr_workspace_bytes(1, 256, 55, 128, _lib.PREC_F16X2)
    assert w16 >= 2 * (n + 9600) * 256 and w32 >= 3 * (w16 - 4096) - 8192
    assert L.sf_corr_workspace_bytes(1, 256, 55, 128, _lib.PREC_FP32_SIMT) >= 2 * 4 * 256 * n
    assert L.sf_corr_workspace_bytes(0, 256, 55, 128, 0) == 0
    assert L.sf_gma_workspace_bytes(3, 128, n, 128) > 3 * n * 128 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_device_means_loud_failure(L):
    from streamflow_b200 import Aggregate, Attention, CorrBlock, StreamCorrError
    assert L.sf_device_ok() != 0
    assert b"no CPU fallback" in L.sf_last_error() or b"no fallback" in L.sf_last_error()
    x = torch.zeros(1, 8, 16, 16)
    with pytest.raises(StreamCorrError):
        CorrBlock(x, x)

    class A:
        pass
    with pytest.raises(StreamCorrError):
        Attention(args=A(), dim=128, heads=1, dim_head=128)(torch.zeros(1, 128, 8, 8))
    with pytest.raises(StreamCorrError):
        Aggregate(args=A(), dim=128, heads=1, dim_head=128)(None, torch.zeros(1, 128, 8, 8))
    # the batched build, the (q, k, fmap) convention and the raw (q, k) entry point refuse CPU tensors the same way
    from streamflow_b200 import CorrGroup
    with pytest.raises(StreamCorrError):
        CorrGroup.from_fmaps(torch.zeros(1, 4, 8, 16, 16))
    with pytest.raises(StreamCorrError):
        CorrGroup.from_fmaps(torch.zeros(1, 1, 8, 16, 16))            # fewer than two frames
    agg = Aggregate(args=A(), dim=128, heads=1, dim_head=128)
    q = torch.zeros(1, 128, 8, 8)
    with pytest.raises(StreamCorrError):
        agg(q, q, q)
    with pytest.raises(TypeError):
        agg(q)
    qk = Attention(args=A(), dim=128, heads=1, dim_head=128, return_qk=True)(q)      # the projection itself is torch
    assert len(qk) == 2 and qk[0].shape == (1, 128, 8, 8)
    assert L.sf_gma_attention_qk(1024, 1024, 0, 1, 64, 128, 0.1, 1, 1024, 1024, 1024, 1 << 30, None) != 0
    # the raw entry points also fail (return code, not a crash)
    import ctypes
    lv = (ctypes.c_void_p * 4)(1024, 1024, 1024, 1024)
    assert L.sf_corr_lookup(lv, 1024, 1024, 1, 16, 16, 4, 4, None) != 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under streamflow_b200/ may import it."""
    pkg = os.path.join(ROOT, "streamflow_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+[\w.]*oracle", src, re.M), f"{fn} imports the oracle"
            assert "torch_port" not in src and "streamflow_oracle" not in src, f"{fn} references the oracle"
