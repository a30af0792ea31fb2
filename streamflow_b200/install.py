"""Make the reference's own import statements resolve to the B200 operators.

The StreamFlow model code imports its hot-path operators as top-level modules with ``core/`` on ``sys.path``:

    from corr import CorrBlock                                  (core/models/streamflow.py:8, raft.py, ...)
    from gma import Attention, Aggregate, ...                   (core/models/streamflow.py:10, core/update.py:4)

``install()`` registers two shim modules under those names so ``core/models`` and ``core/update.py`` run
unmodified on top of libstreamcorr.so.  ``core/update.py`` also imports ablation aggregators
(``SpatioTemporalAggregate``, ``TemporalAggregate``, ``SpatioTemporalAggregate2``, ``TMMAggregate``,
``TemporalAttention``) that the shipped configuration never instantiates; if the reference's ``gma.py`` is
importable (``reference_core=`` path) they are re-exported from it, otherwise they resolve to a class that raises
on construction.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_SAVED: dict[str, object] = {}
_ABLATIONS = ["RelPosEmb", "TemporalAggregate", "SpatioTemporalAggregate", "SpatioTemporalAggregate2",
              "TMMAggregate", "TemporalAttention"]


def _unavailable(name):
    class _Unavailable:  # noqa: D401
        def __init__(self, *a, **k):
            raise NotImplementedError(
                f"gma.{name} is an ablation variant outside the B200 hot path; pass reference_core= to "
                "streamflow_b200.install() to re-export the reference implementation")
    _Unavailable.__name__ = name
    return _Unavailable


def install(reference_core: str | None = None) -> None:
    """Register ``corr`` and ``gma`` shim modules in ``sys.modules`` (idempotent)."""
    from . import corr as _corr
    from . import gma as _gma

    for name in ("corr", "gma"):
        if name not in _SAVED:
            _SAVED[name] = sys.modules.get(name)

    m_corr = types.ModuleType("corr")
    m_corr.__doc__ = "streamflow_b200 shim for the reference's core/corr.py"
    m_corr.CorrBlock = _corr.CorrBlock
    m_corr.CorrGroup = _corr.CorrGroup

    m_gma = types.ModuleType("gma")
    m_gma.__doc__ = "streamflow_b200 shim for the reference's core/gma.py"
    m_gma.Attention = _gma.Attention
    m_gma.Aggregate = _gma.Aggregate
    ref = None
    if reference_core is not None:
        path = os.path.join(reference_core, "gma.py")
        spec = importlib.util.spec_from_file_location("_streamflow_ref_gma", path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    for name in _ABLATIONS:
        setattr(m_gma, name, getattr(ref, name) if ref is not None and hasattr(ref, name) else _unavailable(name))

    sys.modules["corr"] = m_corr
    sys.modules["gma"] = m_gma


def uninstall() -> None:
    for name, old in list(_SAVED.items()):
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old
        _SAVED.pop(name)
