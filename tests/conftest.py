import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads the C-ABI library; CPU-only boxes can still load it."""
    from yoloret_b200.build import build_library
    build_library()
    from yoloret_b200 import _lib
    return _lib.lib()


ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]


@pytest.fixture(scope="session")
def anchors():
    import numpy as np
    return np.array(ANCHORS, np.float32).reshape(-1, 2)
