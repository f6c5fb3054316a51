import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _ensure_built():
    lib = os.path.join(ROOT, "picha_b200", "libpicha_b200.so")
    orc = os.path.join(ROOT, "oracle", "libpicha_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()


_ensure_built()

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def fixtures():
    return np.load(os.path.join(GOLDEN, "picha_fixtures.npz"))


@pytest.fixture(scope="session")
def ref_vectors():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def gpu():
    import picha_b200 as P
    if P.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is usable (picha_b200 has no CPU fallback)")
    return P
