import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_modules():
    import torch
    return torch.load(os.path.join(GOLDEN, "modules.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_loss_boundary():
    import torch
    return torch.load(os.path.join(GOLDEN, "loss_boundary.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_loss():
    import torch
    return torch.load(os.path.join(GOLDEN, "loss.pt"), weights_only=False)
