import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rel_err(a, b):
    """(||a-b||_2 / ||b||_2, max|a-b| / max|b|) -- the two per-tensor measures of SURVEY 8c."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    mb = np.max(np.abs(b)) if b.size else 0.0
    return (np.linalg.norm(a - b) / (nb if nb > 0 else 1.0), np.max(np.abs(a - b)) / (mb if mb > 0 else 1.0))


def assert_close(a, b, tol, what=""):
    e2, em = rel_err(a, b)
    assert e2 <= tol and em <= tol, f"{what}: rel-l2 {e2:.3e}, rel-max {em:.3e} > {tol:.1e}"


@pytest.fixture(scope="session")
def lib():
    from novel_vqa_b200 import _lib
    return _lib.load()
