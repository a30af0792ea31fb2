"""streamflow_b200 -- B200 (sm_100a) implementation of StreamFlow's correlation / GMA hot path.

Public surface mirrors the reference operators (``core/corr.py``, ``core/gma.py``):
``CorrBlock``, ``Attention``, ``Aggregate`` (+ ``CorrGroup`` for pair-batched lookups and
``install()`` to make ``from corr import CorrBlock`` / ``from gma import Attention, Aggregate``
resolve to these classes so the unchanged StreamFlow model code runs on them).
"""
from ._lib import StreamCorrError, lib  # noqa: F401
from .corr import CorrBlock, CorrGroup, coords_grid  # noqa: F401
from .gma import Aggregate, Attention, AttentionHandle  # noqa: F401
from .graph import GraphedCall, GraphedModel  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .pcblock import patch_motion_encoder, pcblock_ffn1, unpatch_motion_encoder  # noqa: F401
from .upsample import patch_upsample, upsample_flow  # noqa: F401

__all__ = ["CorrBlock", "CorrGroup", "coords_grid", "Attention", "Aggregate", "AttentionHandle", "install",
           "uninstall", "upsample_flow", "patch_upsample", "StreamCorrError", "lib", "GraphedCall", "GraphedModel",
           "pcblock_ffn1", "patch_motion_encoder", "unpatch_motion_encoder"]
