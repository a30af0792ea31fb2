"""Shared helpers for the test-suite (input generators identical to tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rs_normal(seed, shape):
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


def rs_uniform(seed, lo, hi, shape):
    return np.random.RandomState(seed).uniform(lo, hi, shape).astype(np.float32)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (SURVEY 'Hard parts': the parity norm)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel_err(a, b):
    """max|a-b| / max|b|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def coord_sets(seed, b, h, w):
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    g = np.broadcast_to(np.stack([xs, ys], 0).astype(np.float32)[None], (b, 2, h, w)).copy()
    sets = {
        "grid": g,
        "half": g + 0.5,
        "jitter": g + 3.0 * rs_normal(seed + 10, g.shape),
        "far": g + rs_uniform(seed + 11, -80, 80, g.shape),
        "neg": g - 6.25,
        "border": g * np.float32(1.0) + rs_uniform(seed + 12, -1, 1, g.shape) * np.array(
            [w, h], np.float32).reshape(1, 2, 1, 1),
    }
    return {k: v.astype(np.float32) for k, v in sets.items()}
