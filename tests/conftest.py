import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def qs():
    import quicksand_b200 as q
    from quicksand_b200 import _lib
    _lib.lib()   # fails loudly if libqsb200.so is missing
    return q
