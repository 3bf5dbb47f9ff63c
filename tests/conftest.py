import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def rel_err(a, b):
    """max|a-b| / max|b|  (per-tensor max-norm relative error)."""
    import torch
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    diff = (a - b).abs().max().item() if a.numel() else 0.0
    return diff / max(denom, 1e-30) if denom > 0 else diff


@pytest.fixture(scope='session')
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name), weights_only=False)
    return load
