import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_example():
    return np.load(os.path.join(GOLDEN, "example_seed0.npz"))


@pytest.fixture(scope="session")
def golden_repro():
    return np.load(os.path.join(GOLDEN, "reproduction_seed4.npz"))


@pytest.fixture(scope="session")
def golden_hankel():
    return np.load(os.path.join(GOLDEN, "hankel.npz"))


@pytest.fixture(scope="session")
def refclass():
    """Fixtures produced by the UNMODIFIED reference controller class (tests/golden/make_golden_refclass.py)."""
    return {k: np.load(os.path.join(GOLDEN, f"refclass_{k}.npz")) for k in
            ("example_seed0", "reproduction_seed4", "variants", "errors", "config4", "short_data")}
