import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are SKIPPED (plain `pytest tests/` stays green on a CPU-only machine);
    with one they run, and the product has no CPU path to fall back to."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device: pegasus_b200 has no CPU path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
