import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: no CPU fallback.
    pass


@pytest.fixture(scope="session")
def scenes_dir():
    return os.path.join(ROOT, "scenes")


@pytest.fixture(scope="session")
def small_stars():
    """A dense-ish synthetic catalogue small enough for brute-force checks."""
    from blackstar_b200 import starmap
    return starmap.synthetic_stars(20000, seed=7)
