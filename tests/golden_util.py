import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)
import cases  # noqa: E402


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    """max |a-b| / max(|a|,|b|,floor), floor = 1e-12*max|b|  (SURVEY.md 8(d) parity gates)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-12 * scale)
    return float(np.max(np.abs(a - b) / denom))


def scaled_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)
