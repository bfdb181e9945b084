import os
import sys

import pytest

# one hardware work queue per stream (the default 8 are shared round-robin): the in-process multi-rank tests drive several engines
# from one process, and a barrier kernel that spins at the head of a shared queue would block the peer's kernels queued behind it.
# Has to be in the environment before CUDA initialises; one process per GPU (the product layout) is not affected.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and every kernel loaded up front: with lazy loading the first launch of a kernel waits for the device to drain, which a spinning
# barrier kernel of the other in-process rank never lets happen (two processes do not have this problem: the peer arrives on its own)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

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
