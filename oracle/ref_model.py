"""Load the UNMODIFIED reference model around either set of hot-path operators (TEST INFRASTRUCTURE ONLY).

    mod = load_model_module("reference")   # core/models/streamflow.py on the reference's own corr.py / gma.py
    mod = load_model_module("b200")        # the same file after streamflow_b200.install(): `from corr import
                                           # CorrBlock`, `from gma import Attention, Aggregate` (streamflow.py:8,10,
                                           # update.py:4) resolve to the B200 operators
    model = build_model(mod, T=4)          # SKFlow_MF8 with the shipped configuration

The reference sources come from oracle/_ref (oracle/make_ref.py) or /root/reference; `timm` comes from the real
package if installed, else from oracle/timm_shim.  The two variants are separate module objects (`update` and the
model file are executed once per variant), so both can live in one process and share weights via state_dict.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import os
import sys
import warnings

import torch

from .make_ref import ref_core_dir

HERE = os.path.dirname(os.path.abspath(__file__))
_L1_DEPENDENT = ("corr", "gma", "update")
_LOADED: dict = {}


def available() -> bool:
    return ref_core_dir() is not None


def _ensure_timm():
    try:
        import timm  # noqa: F401
    except ModuleNotFoundError:
        sys.path.insert(0, os.path.join(HERE, "timm_shim"))
        import timm  # noqa: F401


class ShippedArgs(argparse.Namespace):
    """argparse.Namespace of the shipped configuration (scripts/infer.sh:12-26, evaluate_mf.py:1100-1197).

    `SKFlow_MF8.__init__` passes `args` positionally into `Twins_CSC(pretrained=...)` (streamflow.py:45 vs
    twins_csc.py:38): a truthy Namespace makes the encoder torch.load('./pretrained/twins_svt_large-90f6aaa9.pth').
    Random-init runs need the `pretrained=False` path (as test_memory.py:119 does), hence the falsy __bool__."""

    def __bool__(self):
        return False


def shipped_args(T: int = 4, mixed_precision: bool = True) -> ShippedArgs:
    return ShippedArgs(model_name="SKFlow_MF8", Encoder="Twins_CSC", UpdateBlock="SKUpdateBlock_TAM_v3",
                       MotionEncoder="SKMotionEncoder6_Deep_nopool_res", decoder_dim=256, num_heads=1, use_gma=True,
                       k_conv=[1, 15], PCUpdater_conv=[1, 7], T=T, mixed_precision=mixed_precision, dropout=0,
                       corr_levels=4, corr_radius=4)


def load_model_module(l1: str):
    """Execute core/models/streamflow.py (and core/update.py under it) with `corr` / `gma` bound to `l1`."""
    if l1 not in ("reference", "b200"):
        raise ValueError(l1)
    if l1 in _LOADED:
        return _LOADED[l1]
    core = ref_core_dir()
    if core is None:
        raise RuntimeError("reference sources not available: run `python oracle/make_ref.py` where /root/reference exists")
    _ensure_timm()
    if core not in sys.path:
        sys.path.insert(0, core)
    saved = {k: sys.modules.pop(k, None) for k in _L1_DEPENDENT}
    try:
        if l1 == "b200":
            import streamflow_b200
            streamflow_b200.install(reference_core=core)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = importlib.util.spec_from_file_location(f"_streamflow_model_{l1}",
                                                          os.path.join(core, "models", "streamflow.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        mod._l1_modules = {k: sys.modules.get(k) for k in _L1_DEPENDENT}
    finally:
        if l1 == "b200":
            import streamflow_b200
            streamflow_b200.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _LOADED[l1] = mod
    return mod


def build_model(mod, T: int = 4, mixed_precision: bool = True):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mod.SKFlow_MF8(shipped_args(T, mixed_precision))


@torch.no_grad()
def randomise(model, seed: int = 0, flow_gain: float = 1.0):
    """Seeded re-initialisation of the two blocks the reference zero-initialises: gamma ~ U(0.5,
    1.5) (core/gma.py:84) and TemporalLayer2's transformer block (core/update.py:453-457,505: trunc_normal std .02
    weights, LayerNorm weight 1) -- otherwise GMA and the temporal path contribute nothing at init.  `flow_gain`
    scales the last conv of the flow head so that 12 iterations produce Sintel-like magnitudes."""
    g = torch.Generator().manual_seed(seed)
    ub = model.update_block
    ub.aggregator.gamma.copy_(0.5 + torch.rand(1, generator=g))
    for name, p in ub.transformer_block.named_parameters():
        if "norm" in name:
            p.fill_(1.0 if name.endswith("weight") else 0.0)
        elif p.dim() > 1:
            w = torch.empty(p.shape)
            torch.nn.init.trunc_normal_(w, std=0.02, generator=g)
            p.copy_(w)
        else:
            p.zero_()
    last = ub.flow_head.ffn2[2]
    last.weight.mul_(flow_gain)
    last.bias.mul_(flow_gain)
    return model
