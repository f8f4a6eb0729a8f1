import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def rel_err(a, b):
    """(max|a-b| / max|b|, ||a-b||_2 / ||b||_2) — the parity metric of SURVEY §8(d)."""
    import torch
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    d = (a - b)
    return float(d.abs().max() / b.abs().max().clamp_min(1e-30)), float(d.norm() / b.norm().clamp_min(1e-30))
