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
    # The parts of the model that are still torch library calls (cuDNN convs of the EfficientViMBlock shell) must not
    # silently drop to TF32 while we check a 1e-4 parity gate.
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/*.npz fixture (see tests/golden/make_golden.py)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def t(self, key, device="cpu"):
        return torch.from_numpy(self.z[key]).to(device)

    def has(self, key):
        return key in self.z.files

    def sd(self, device="cpu", after=False):
        pre = "sd_after/" if after else "sd/"
        return {k[len(pre):]: torch.from_numpy(self.z[k]).to(device) for k in self.z.files if k.startswith(pre)}

    def grads(self, device="cpu"):
        return {k[5:]: torch.from_numpy(self.z[k]).to(device) for k in self.z.files if k.startswith("grad/")}


@pytest.fixture
def golden():
    return Golden


def rel_err(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the tolerance quoted in north_star)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
