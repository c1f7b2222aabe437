"""Shared pytest configuration: the ``gpu`` marker and fixture loaders."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped automatically when no CUDA device is visible."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden_grid_ic():
    return load_golden("grid_ic.npz")


@pytest.fixture(scope="session")
def golden_single_step():
    return load_golden("single_step.npz")


@pytest.fixture(scope="session")
def golden_multi_step():
    return load_golden("multi_step.npz")


@pytest.fixture(scope="session")
def golden_ref10():
    return load_golden("ref_data_10.npz")


def normwise(a, b):
    """max|a-b| / max|b| -- the per-field parity measure."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
