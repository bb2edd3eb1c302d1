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
    """Golden .npz -> (HeadCase or None, dict of arrays)."""
    from odcp_b200 import synthetic, targets
    z = dict(np.load(os.path.join(GOLDEN, name)))
    case = None
    if "rec" in z:
        rec = np.ascontiguousarray(z["rec"]).reshape(-1).view(targets.GT_DTYPE).copy()
        y = torch.from_numpy(z["y"]) if "y" in z else None
        case = synthetic.HeadCase(name, int(z["version"]), int(z["n"]), int(z["s_h"]), int(z["s_w"]),
                                  int(z["a"]), int(z["c"]), int(z["height"]), int(z["width"]), y, rec,
                                  z["gt_off"].astype(np.int32))
    return case, z


def golden_lambdas(z):
    from odcp_b200 import synthetic
    if "lambdas" in z:
        keys = ("lambda_xy", "lambda_wh", "lambda_conf", "lambda_noobj", "lambda_cls")
        return {k: float(v) for k, v in zip(keys, z["lambdas"])}
    return dict(synthetic.DEFAULT_LAMBDAS)


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative' of the north-star tolerance for tensors."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
