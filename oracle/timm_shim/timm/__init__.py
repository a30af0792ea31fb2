"""Minimal stand-in for the `timm` package (TEST INFRASTRUCTURE ONLY -- never imported by streamflow_b200).

The reference model (core/models/streamflow.py, core/update.py, core/encoders/*.py) depends on timm, which is
un-vendored and unpinned (`install.sh:2`: `pip install ... timm ...`; era torch 1.13.1 => timm 0.9.x) and is not
installed in this image (no wheel, no network).  This shim restates exactly the pieces the reference touches so
the UNMODIFIED reference model can run around the hot-path operators for the end-to-end EPE check:

  timm.create_model('twins_svt_large', pretrained=False)     core/encoders/twins_csc.py:40
  timm.models.twins.{GlobalSubSampleAttn, LocallyGroupedAttn} core/encoders/twins_csc.py:8
  timm.models.vision_transformer.Attention                    core/update.py:450,466; core/models/streamflow.py:4
  timm.models.layers / timm.layers: Mlp, DropPath, to_2tuple, drop_path, trunc_normal_
  timm.models.registry.register_model                         core/encoders/umt.py:7

Hyper-parameters of twins_svt_large follow the reference's own restatement at core/encoders/twins_1dconv.py:58-75
(depths [2,2,18,2], dims [128,256,512,1024], heads [4,8,16,32], mlp_ratio 4, sr_ratios [8,4,2,1], window 7,
LayerNorm eps 1e-6; block i is window attention if i is even, global sub-sampled otherwise) and its Block wiring at
:15-31.  Module / parameter names follow timm 0.9 so that reference checkpoints would load with strict=True.
"Parity unpinned" with respect to real timm: nothing in the reference pins encoder outputs; the end-to-end test
compares reference-L1 vs B200-L1 under the SAME surrounding code, so the shim only has to be a faithful caller.
"""
from . import layers, models  # noqa: F401
from .models.twins import create_model  # noqa: F401

__version__ = "0.9.shim"
