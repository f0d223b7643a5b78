import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def mini_cfg():
    from oracle.birefnet_ref import Config
    return Config.mini()


@pytest.fixture(scope="session")
def mini_weights_A(mini_cfg):
    from oracle.make_weights import make_weights
    return make_weights(mini_cfg, seed=0, weight_set="A")


@pytest.fixture(scope="session")
def mini_weights_B(mini_cfg):
    from oracle.make_weights import make_weights
    return make_weights(mini_cfg, seed=0, weight_set="B", offset_sigma=2.0)
