"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 via gpurun."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_ref: needs oracle/_ref/libheston_ref.so (prebuilt in the build container)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own heston.cpp compiled unmodified; skipped if neither the prebuilt .so
    nor /root/reference is present."""
    from oracle.oracle import Reference

    try:
        return Reference()
    except (FileNotFoundError, OSError) as e:  # pragma: no cover
        pytest.skip(str(e))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def g_cf():
    return load_golden("ref_cf.npz")


@pytest.fixture(scope="session")
def g_prices():
    return load_golden("ref_prices.npz")


@pytest.fixture(scope="session")
def g_misc():
    return load_golden("ref_misc.npz")


@pytest.fixture(scope="session")
def g_cal():
    return load_golden("ref_calibrator.npz")


@pytest.fixture(scope="session")
def g_sabr():
    return load_golden("ref_sabr.npz")


@pytest.fixture(scope="session")
def g_fft():
    return load_golden("fft_selfcheck.npz")


@pytest.fixture(scope="session")
def g_fft_direct():
    return load_golden("fft_direct.npz")


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a)
    b = np.asarray(b)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
