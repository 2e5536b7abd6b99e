import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def leafoff_state():
    from oracle import weights
    weights.export_reference_checkpoints()
    return weights.load_state("leafoff")


@pytest.fixture(scope="session")
def random_state():
    from oracle.network import random_state as rs
    return rs(seed=7)


@pytest.fixture(scope="session")
def small_tree():
    from smartqsm_b200.synthetic import make_tree
    return make_tree(30000, 11)
