import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def mcu_built():
    """Path of the in-tree libmambacuda.so (built by mamba.jl_b200/build.py if sources are newer)."""
    sys.path.insert(0, os.path.join(ROOT, "mamba.jl_b200"))
    import build as mcu_build
    return mcu_build.build()
