"""Small modules of the reference's cliora/net/utils.py that callers import by name."""
import torch
import torch.nn as nn

from .index import Index, get_inside_index, get_offset_cache, get_outside_index  # noqa: F401

TINY = 1e-8


class UnitNorm(object):
    """x / max(||x||, 1e-8)  (cliora/net/utils.py:11-14).  Host-side helper for tensors outside the chart."""

    def __call__(self, x, p=2, eps=TINY):
        return x / x.norm(p=p, dim=-1, keepdim=True).clamp(min=eps)


class NormalizeFunc(nn.Module):
    def __init__(self, mode='none'):
        super().__init__()
        self.mode = mode

    def forward(self, x):
        return UnitNorm()(x) if self.mode == 'unit' else x


class BatchInfo(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


def __getattr__(name):
    # the reference keeps ImageEncoder in cliora/net/utils.py:37-55; here it lives next to the fused GEMM wrappers
    if name == 'ImageEncoder':
        from .trainer import ImageEncoder
        return ImageEncoder
    raise AttributeError(name)
