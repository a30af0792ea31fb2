"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` covers the oracle against the golden vectors, host logic and that the
C-ABI library loads and exports its symbols; ``-m gpu`` is the CUDA parity suite.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _inference_only():
    """The B200 operators are inference-only and refuse to run where autograd would expect a graph (parameters that
    require grad with grad mode on); the reference drives them under torch.no_grad() (evaluate_mf.py:468).  Tests
    that check the refusal itself re-enable grad locally."""
    import torch
    with torch.no_grad():
        yield
