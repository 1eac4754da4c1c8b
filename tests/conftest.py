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


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack_gt(gt, count):
    return [torch.from_numpy(gt[i, : int(count[i])].copy()) for i in range(gt.shape[0])]


def golden_loss_inputs(name):
    """(fixture, preds, gts, anchors, strides, grad) of a stored loss case, in its stored dtype."""
    z = load_golden(name)
    bf16 = bool(z["meta"][5])
    conv = (lambda a: torch.from_numpy(a.copy()).view(torch.bfloat16)) if bf16 else (lambda a: torch.from_numpy(a.copy()))
    return (z, conv(z["preds"]), unpack_gt(z["gt"], z["gt_count"]), conv(z["anchors"]), conv(z["strides"]),
            conv(z["grad"]))


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
