import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (cheap no-op when they are up to date).  On the GPU
    box the prebuilt .so files travel with the snapshot; nvcc/gcc are only needed if they are stale."""
    from mach3_b200 import build
    build.build_synth()
    build.build_oracle()
    build.build_cuda()
    yield
