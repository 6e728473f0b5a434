import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs /root/reference (this container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def world_big():
    from world import World
    return World.load_npz(os.path.join(GOLDEN, "world_big.npz"))


@pytest.fixture(scope="session")
def world_small():
    from world import World
    return World.load_npz(os.path.join(GOLDEN, "world_small.npz"))
