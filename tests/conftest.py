import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "csi-nn2_b200", "lib", "libb200nn.so"))
        return lib.b200_device_count() > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the b200 path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from shl import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from shl import Harness
    return Harness("ref")


@pytest.fixture(scope="session")
def ref_noavx():
    from shl import Harness
    return Harness("ref_noavx")


@pytest.fixture(scope="session")
def b200():
    from shl import Harness
    return Harness("b200")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "unit_kat.npz")))


@pytest.fixture()
def rng():
    return np.random.default_rng(1234)
