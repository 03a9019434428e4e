import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/liboracle.so), built on demand with the committed recipe."""
    from oracle import binding as ob
    ob.build_oracle()
    return ob.Oracle()


@pytest.fixture(scope="session")
def lib():
    """The product's C-ABI library. Built on demand (nvcc cross-compiles without a GPU)."""
    from brickmap_b200 import build as b
    b.build()
    import brickmap_b200
    return brickmap_b200.load()


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, "golden_%s.npz" % name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get
