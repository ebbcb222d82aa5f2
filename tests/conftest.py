import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def weights():
    from oracle import mvsnerf_oracle as orc
    return orc.load_weights_npz(os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))


def load_golden(name):
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, name))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.dtype != object else a
    return out


@pytest.fixture(scope="session")
def golden_tiny():
    return load_golden("tiny_32x32_pad4.npz")


@pytest.fixture(scope="session")
def golden_tiny_lindisp():
    return load_golden("tiny_32x32_pad0_lindisp.npz")


@pytest.fixture(scope="session")
def golden_c1():
    return load_golden("c1_64x64_pad24.npz")


@pytest.fixture(scope="session")
def golden_bn_modes():
    return load_golden("bn_modes_64x96_pad4.npz")


@pytest.fixture(scope="session")
def golden_grad():
    return load_golden("grad_tiny_32x32_pad4.npz")
