import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# the reference's own test files are fixtures, run through tests/test_reference_unit_tests.py
collect_ignore_glob = ["golden/ref_tests/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the /root/reference checkout")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_close(a, b, rtol=1e-12, atol=0.0, msg=""):
    """Relative comparison that treats matching inf/nan as equal."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s shape %s vs %s" % (msg, a.shape, b.shape)
    fin = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), msg + " nan pattern"
    assert np.array_equal(a[~fin & ~np.isnan(a)], b[~fin & ~np.isnan(b)]), msg + " inf pattern"
    err = np.abs(a[fin] - b[fin])
    tol = atol + rtol * np.maximum(np.abs(a[fin]), np.abs(b[fin]))
    bad = err > tol
    if np.any(bad):
        i = np.argmax(err - tol)
        raise AssertionError("%s: %d/%d outside rtol=%g; worst |%r - %r| = %g" % (
            msg, bad.sum(), bad.size, rtol, a[fin][i], b[fin][i], err[i]))


@pytest.fixture(autouse=True)
def _default_rng_mode():
    """Every test starts in the product's default draw mode (tests that switch modes need no clean-up)."""
    try:
        import hepmc_b200 as H
        H.rng.set_mode(H.rng.DEFAULT_MODE)
    except Exception:
        pass
    yield
