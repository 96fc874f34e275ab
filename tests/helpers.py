"""Shared comparison helpers for the parity tests."""
import numpy as np


def rel_err(a, b):
    """max|a-b| / max|b|  — the parity metric of SURVEY.md §8(d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / den


def assert_close(a, b, rtol, what=""):
    e = rel_err(a, b)
    assert e <= rtol, "%s: rel err %.3e > %.1e" % (what, e, rtol)
    return e
