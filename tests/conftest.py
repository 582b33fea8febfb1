import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name in ("grid_edges", "builders", "model", "resize", "mlp"):
        with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
            out[name] = {k: z[k] for k in z.files}
    return out


@pytest.fixture(scope="session")
def libgnc():
    """Builds (if stale) and loads libgnc.so."""
    from graphnet_classifier_b200 import _lib, build
    build.build()
    return _lib.load()
