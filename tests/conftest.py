import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine_lib():
    """Build (if stale) and load the C-ABI engine; never skipped -- a missing library is a failure."""
    from gps_slam_b200 import build, engine
    build.build()
    return engine.load_library()
