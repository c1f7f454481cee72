import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def device():
    """One GeometryDevice (t3d_ctx) for the whole GPU test session; creation fails loudly without a GPU."""
    import tetra3d_b200 as t3d
    return t3d.GeometryDevice(0)
