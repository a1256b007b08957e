import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import _oracle
    return _oracle.lib()


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library; GPU tests fail loudly (no skip, no CPU fallback) if it is missing or no device works."""
    import armour_b200
    return armour_b200.lib()
