import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (compiled in-tree on first use; nvcc cross-compiles without a GPU)."""
    from graphneuralnetwork_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build()
    return _lib.load()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b):
    """max |a-b| / max |b| — the relative tolerance north_star states (1e-5 fp32, 1e-2 bf16)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / denom)


def assert_close_elementwise(a, b, rtol=1e-4, atol_rms=1e-5):
    """Element-wise companion of rel_err (max-norm alone leaves small-magnitude outputs unchecked):
    |a - b| <= rtol*|b| + atol_rms*rms(row of b) for EVERY element (the rms of the element's own row: an fp32
    sum's rounding error scales with the magnitude of what the row adds, and hub rows add thousands of terms)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    b2 = b.reshape(b.shape[0], -1) if b.ndim > 1 else b.reshape(1, -1)
    rms = np.sqrt(np.mean(b2 * b2, axis=1, keepdims=True)) if b.size else np.zeros((1, 1))
    tol = rtol * np.abs(b2) + atol_rms * np.maximum(rms, 1e-30)
    bad = np.abs(a.reshape(b2.shape) - b2) > tol
    assert not bad.any(), f"{int(bad.sum())} of {b.size} elements off; worst |diff| {np.abs(a.reshape(b2.shape) - b2).max():.3e}"
