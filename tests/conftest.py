import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def _load(name):
        return dict(np.load(golden_path(name)))

    return _load


@pytest.fixture(autouse=True)
def _restore_library_modes(request):
    """Tests that switch the library's schedule overrides (config.set_potrf_mode / set_trsv_mode) leave them at auto."""
    yield
    if request.node.get_closest_marker("gpu") is not None:
        from pygpr_b200 import config
        config.set_potrf_mode("auto")
        config.set_trsv_mode("auto")
        config.set_ozaki_bk(0)
