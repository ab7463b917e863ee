import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """tests/golden/<name>.npz as a dict of torch tensors (weights keyed like the state_dict)."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k.replace("__", "."): torch.from_numpy(z[k].copy()) for k in z.files}


def hot_params(g):
    return {k: v for k, v in g.items() if k == "alpha" or k.split(".")[0] in ("imnet", "flow_imnet", "synth_net")}


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(1.0 / mse)
