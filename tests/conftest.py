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
    """The CPU oracle (oracle/oracle.cpp) -- the checker, never the thing under test."""
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(autouse=True)
def _reset_oracle_modes():
    yield
    import oracle as orc
    orc.set_modes(orc.MATH_NATIVE, orc.RNG_PCG3D)
