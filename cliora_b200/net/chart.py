"""Chart container + the autograd bridge from torch to the C-ABI chart kernels.

Replaces the per-level Python loops of the reference (cliora/net/diora.py:295-398,
cliora/net/cliora.py:304-414) with four library calls: inside/outside forward and
outside/inside backward.  torch owns every buffer; the library only sees pointers.
"""
import ctypes

import torch

from .. import _lib
from .._lib import Dims, WeightGrads, Weights, check, ptr

_SHARED_KEYS = ['W_leaf', 'b_leaf', 'W1', 'b1', 'W2', 'b2', 'Wb', 'root']
_OUTSIDE_KEYS = ['oW1', 'ob1', 'oW2', 'ob2', 'oWb']


class Chart(object):
    """Same attributes as the reference Chart (diora.py:7-23).  The ``*_c`` tensors are identically
    zero in the MLP architecture (diora.py:70) and are only materialised if somebody reads them."""

    def __init__(self, inside_h, inside_s, outside_h, outside_s):
        self.inside_h, self.inside_s = inside_h, inside_s
        self.outside_h, self.outside_s = outside_h, outside_s
        self._inside_c = self._outside_c = self._vis = None

    @property
    def inside_c(self):
        if self._inside_c is None:
            self._inside_c = torch.zeros_like(self.inside_h)
        return self._inside_c

    @property
    def outside_c(self):
        if self._outside_c is None:
            self._outside_c = torch.zeros_like(self.outside_h)
        return self._outside_c

    @property
    def vis_aggragate(self):  # sic, cliora.py:25 (allocated by the reference, never written)
        if self._vis is None:
            self._vis = torch.zeros_like(self.inside_h)
        return self._vis


class ChartRun(object):
    """Per-forward bookkeeping the module needs after the kernels ran (workspace views for hooks, CKY)."""

    def __init__(self):
        self.ws = None
        self.layout = None
        self.B = self.n = self.D = self.R = None
        self.consumed = False

    def level_rows(self, level, outside=False):
        L = _lib.lib()
        r0 = int(L.cliora_split_row_offset(self.B, self.n, level, 1 if outside else 0))
        Lc = self.n - level
        N = (self.n - level - 1) if outside else level
        return r0, self.B * Lc * N

    def split_h(self, level, outside=False):
        """Pre-aggregation vectors of a level, [B*L*N, D] -- the ``h`` of inside_hook (diora.py:331)."""
        r0, rows = self.level_rows(level, outside)
        base = self.layout.Yout if outside else self.layout.Yin
        D = self.D
        return self.ws[base + r0 * D: base + (r0 + rows) * D].view(rows, D)

    def split_z(self, level, outside=False):
        """Hidden activations relu(W1 [l;r] + b1) of a level, [rows, D] (tf32-rounded part when the
        tensor-core path stores split pairs; its sign pattern is exact)."""
        r0, rows = self.level_rows(level, outside)
        base = self.layout.Zout if outside else self.layout.Zin
        D = self.D
        return self.ws[base + r0 * D: base + (r0 + rows) * D].view(rows, D)

    def split_s(self, level, outside=False):
        """Raw split scores of a level, [B,L,N,1] inside / [B,N,L,1] outside -- the ``s`` of the hooks."""
        r0, rows = self.level_rows(level, outside)
        base = self.layout.Eout if outside else self.layout.Ein
        s = self.ws[base + r0: base + r0 + rows]
        Lc = self.n - level
        return s.view(self.B, -1, Lc, 1) if outside else s.view(self.B, Lc, -1, 1)

    def all_split_scores(self):
        """The whole inside split-score region (input of the CKY kernel)."""
        return self.ws[self.layout.Ein: self.layout.Ein + max(int(self.layout.rows_in), 1)]


def _weights_struct(cls, tensors, share):
    s = cls()
    keys = _SHARED_KEYS + ([] if share else _OUTSIDE_KEYS)
    for k, t in zip(keys, tensors):
        setattr(s, k, ptr(t))
    return s


class ChartFunction(torch.autograd.Function):
    """(x, obj, weights) -> (inside_h, inside_s, outside_h, outside_s) with a hand-written backward."""

    @staticmethod
    def forward(ctx, run, share, outside, x, obj, keep, *weights):
        L = _lib.lib()
        if not x.is_cuda:
            raise _lib.ClioraError('cliora_b200: the chart runs on CUDA only (input is on %s); no CPU fallback'
                                   % x.device)
        x = x.detach().contiguous().float()
        obj = None if obj is None else obj.detach().contiguous().float()
        keep = None if keep is None else keep.detach().contiguous().to(torch.uint8)
        weights = tuple(w.detach().contiguous() for w in weights)
        B, n, D = x.shape
        R = 0 if obj is None else obj.shape[1]
        C = n * (n + 1) // 2
        lay = _lib.layout(B, n, D, R, share)
        dev = x.device
        with torch.cuda.device(dev):
            ws = torch.empty(int(lay.ws_floats), device=dev, dtype=torch.float32)
            inside_h = torch.empty(B, C, D, device=dev, dtype=torch.float32)
            inside_s = torch.empty(B, C, 1, device=dev, dtype=torch.float32)
            if outside:
                outside_h = torch.empty(B, C, D, device=dev, dtype=torch.float32)
                outside_s = torch.empty(B, C, 1, device=dev, dtype=torch.float32)
            else:  # the reference leaves the outside chart at its zero fill (diora.py:19-22)
                outside_h = torch.zeros(B, C, D, device=dev, dtype=torch.float32)
                outside_s = torch.zeros(B, C, 1, device=dev, dtype=torch.float32)
            dims = Dims(B, n, D, R, 1 if share else 0, 0)
            W = _weights_struct(Weights, weights, share)
            st = _lib.stream()
            check(L.cliora_inside_fwd(ctypes.byref(dims), ctypes.byref(W), ptr(x), ptr(obj), ptr(keep),
                                      ptr(inside_h), ptr(inside_s), ptr(ws), st), 'cliora_inside_fwd')
            if outside:
                check(L.cliora_outside_fwd(ctypes.byref(dims), ctypes.byref(W), ptr(inside_h), ptr(inside_s),
                                           ptr(outside_h), ptr(outside_s), ptr(ws), st), 'cliora_outside_fwd')
        run.ws, run.layout, run.B, run.n, run.D, run.R = ws, lay, B, n, D, R
        ctx.run, ctx.share, ctx.outside, ctx.dims_t = run, share, outside, (B, n, D, R)
        ctx.has_obj, ctx.has_keep = obj is not None, keep is not None
        saved = [x, ws, inside_h, inside_s, outside_h, outside_s]
        if obj is not None:
            saved.append(obj)
        if keep is not None:
            saved.append(keep)
        ctx.save_for_backward(*saved, *weights)
        return inside_h, inside_s, outside_h, outside_s

    @staticmethod
    def backward(ctx, g_ih, g_is, g_oh, g_os):
        L = _lib.lib()
        if ctx.run.consumed:
            raise _lib.ClioraError('cliora_b200: the chart backward consumes its workspace; '
                                   'backward through the same forward twice is not supported')
        saved = list(ctx.saved_tensors)
        x, ws, inside_h, inside_s, outside_h, outside_s = saved[:6]
        i = 6
        obj = keep = None
        if ctx.has_obj:
            obj = saved[i]; i += 1
        if ctx.has_keep:
            keep = saved[i]; i += 1
        weights = saved[i:]
        B, n, D, R = ctx.dims_t
        share, outside = ctx.share, ctx.outside
        lay = ctx.run.layout
        dev = x.device
        cont = lambda g: None if g is None else g.contiguous().float()
        g_ih, g_is, g_oh, g_os = cont(g_ih), cont(g_is), cont(g_oh), cont(g_os)
        if not outside:
            g_oh = g_os = None
        with torch.cuda.device(dev):
            bws = torch.empty(int(lay.bws_floats), device=dev, dtype=torch.float32)
            grads = [torch.empty_like(w) for w in weights]
            gx = torch.empty_like(x)
            gobj = torch.empty_like(obj) if obj is not None else None
            dims = Dims(B, n, D, R, 1 if share else 0, 0)
            W = _weights_struct(Weights, weights, share)
            G = _weights_struct(WeightGrads, grads, share)
            st = _lib.stream()
            check(L.cliora_chart_bwd_begin(ctypes.byref(dims), ptr(g_ih), ptr(g_is), ptr(g_oh), ptr(g_os),
                                           ptr(bws), st), 'cliora_chart_bwd_begin')
            if outside:
                check(L.cliora_outside_bwd(ctypes.byref(dims), ctypes.byref(W), ptr(inside_h), ptr(inside_s),
                                           ptr(outside_h), ptr(outside_s), ptr(ws), ptr(bws), ctypes.byref(G), st),
                      'cliora_outside_bwd')
            check(L.cliora_inside_bwd(ctypes.byref(dims), ctypes.byref(W), ptr(x), ptr(obj), ptr(keep),
                                      ptr(inside_h), ptr(inside_s), ptr(outside_h), ptr(ws), ptr(bws),
                                      1 if outside else 0, ptr(gx), ptr(gobj), ctypes.byref(G), st),
                  'cliora_inside_bwd')
        ctx.run.consumed = True
        return (None, None, None, gx, gobj, None) + tuple(grads)
