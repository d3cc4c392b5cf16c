import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", name + ".pt")


@pytest.fixture(scope="session")
def goldens():
    import torch
    out = {}
    for n in ("id_small_collide", "id_cfg1_shape", "text_tiny", "text_tiny_collide"):
        out[n] = torch.load(golden_path(n), map_location="cpu", weights_only=False)
    return out
