import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def small_mesh():
    from eutropia_b200 import mesh
    return mesh.make_mesh("Rectangle_30_24", 1.0, 3)


@pytest.fixture(scope="session")
def harbor_mesh():
    from eutropia_b200 import mesh
    return mesh.make_mesh(mesh.harbor_polygon(70.0), 1.0, 4)
