"""Chart geometry.  Mirrors ``Index`` of the reference (cliora/net/utils.py:67-134) for callers
that still ask for index tensors (scripts/parse.py:245, phrase_embed.py:79-92); the kernels
themselves use the closed forms inline and never read these tensors."""
import ctypes

import torch

from .. import _lib


def get_offset_cache(length):
    """cliora/net/offset_cache.py:1-7 -> {level: first cell index of that level}."""
    L = _lib.lib()
    return {lvl: int(L.cliora_level_offset(length, lvl)) for lvl in range(length)}


def _pair(fn, count, *args):
    a = (ctypes.c_int64 * count)()
    b = (ctypes.c_int64 * count)()
    _lib.check(fn(*args, a, b), fn.__name__)
    return torch.tensor(list(a), dtype=torch.int64), torch.tensor(list(b), dtype=torch.int64)


def get_inside_index(length, level, offset_cache=None, cuda=False):
    """cliora/net/inside_index.py:182-197: (left, right) child indices, flattened (pos, split)."""
    l, r = _pair(_lib.lib().cliora_inside_index, (length - level) * level, length, level)
    return (l.cuda(), r.cuda()) if cuda else (l, r)


def get_outside_index(length, level, offset_cache=None, cuda=False):
    """cliora/net/outside_index.py:93-127: (parent, sibling) indices, flattened (split, pos)."""
    p, s = _pair(_lib.lib().cliora_outside_index, (length - level - 1) * (length - level), length, level)
    return (p.cuda(), s.cuda()) if cuda else (p, s)


class Index(object):
    def __init__(self, cuda=False, enable_caching=True):
        self.cuda = cuda
        self.enable_caching = enable_caching
        self.cache = {}

    def _memo(self, key, fn):
        if not self.enable_caching:
            return fn()
        if key not in self.cache:
            self.cache[key] = fn()
        return self.cache[key]

    def get_offset(self, length):
        return self._memo(('offset', length), lambda: get_offset_cache(length))

    def get_inside_index(self, length, level):
        return self._memo(('inside', length, level), lambda: get_inside_index(length, level, cuda=self.cuda))

    def get_outside_index(self, length, level):
        return self._memo(('outside', length, level), lambda: get_outside_index(length, level, cuda=self.cuda))
