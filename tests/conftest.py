"""Shared pytest plumbing: the ``gpu`` marker, repo-root imports and the package handle.

The package directory is ``3dfacerecon_b200`` (the reference's name); it is not a valid Python
identifier, so everything imports it through ``importlib`` (see ``fr()``).
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fr(sub=None):
    name = "3dfacerecon_b200" + ("." + sub if sub else "")
    return importlib.import_module(name)


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def render_golden():
    blob = np.load(os.path.join(GOLDEN, "render_cases.npz"))
    cases = {}
    for key in blob.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = blob[key]
    return cases


@pytest.fixture(scope="session")
def small_model():
    return fr("synth").make_synthetic_model(grid=(23, 31), ndim_shape=12, ndim_exp=5, seed=3, jitter=0.2)
