import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Every test gets a wall-clock limit (pytest-timeout, when installed): a hung test must cost minutes, not a GPU
    call's whole time limit."""
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(600))


@pytest.fixture(scope="session")
def graph_golden():
    return dict(np.load(os.path.join(GOLDEN, "graph_small.npz")))


@pytest.fixture(scope="session")
def noise_golden():
    return dict(np.load(os.path.join(GOLDEN, "noise_small.npz")))


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def pairwise_golden():
    return dict(np.load(os.path.join(GOLDEN, "graph_pairwise.npz")))
