import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the CUDA library; GPU tests fail loudly when it is missing."""
    from pyresample_b200 import _cabi
    if _cabi.needs_rebuild():
        _cabi.build()
    return _cabi.lib()
