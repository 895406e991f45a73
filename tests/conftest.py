import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_python.npz"))


@pytest.fixture(autouse=True)
def _exact_numerics_by_default():
    """The parity tests pin the EXACT compositing numerics (images bit-identical to the CPU oracle); the product's
    default is "fast" (MUFU exp), which the tests that name it compare within the reference tolerances."""
    from pegasus_b200 import _lib
    old = _lib._default_numerics
    _lib.set_numerics("exact")
    yield
    _lib._default_numerics = old
