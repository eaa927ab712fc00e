import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    arrays = dict(np.load(os.path.join(g, "unetpp_golden.npz")))
    with open(os.path.join(g, "unetpp_golden.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def variants_golden():
    """Reference outputs for the constructor-flag variants and the extra optimizers (oracle/make_golden_variants.py)."""
    import json
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    arrays = dict(np.load(os.path.join(g, "unetpp_variants.npz")))
    with open(os.path.join(g, "unetpp_variants.json")) as f:
        meta = json.load(f)
    return arrays, meta
