import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box through gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gd = os.path.join(ROOT, "tests", "golden")
    return {name: np.load(os.path.join(gd, name + ".npz")) for name in
            ("planning", "stages", "closed_loop_v2", "closed_loop_v3", "closed_loop_variants", "actual_trajectory", "rrt")}


@pytest.fixture(scope="session")
def cuda():
    """The CUDA path or nothing: GPU tests must never pass on a fallback."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from uav_ac_b200 import _native as nat
    nat.lib()
    sm, major, minor = nat.device_info(0)
    assert major >= 10, f"sm_100a library on compute capability {major}.{minor}"
    return torch.device("cuda", 0)
