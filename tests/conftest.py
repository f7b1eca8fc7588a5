import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def golden_tiny():
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, "tiny.npz"))


@pytest.fixture(scope="session")
def golden_kitti():
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, "kitti_uncond.npz"))


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) the in-tree shared library once per session."""
    from lidar_layout_b200 import build
    return build.build()
