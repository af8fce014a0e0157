import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """The in-tree native library; built here if missing (nvcc cross-compiles without a GPU)."""
    from opensot_b200 import build, engine
    if not os.path.exists(engine.lib_path()):
        build.build_native()
    return engine.lib()


@pytest.fixture(scope="session")
def port():
    from oracle import portlib
    portlib.lib()
    return portlib
