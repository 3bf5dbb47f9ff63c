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
    """Per-forward bookkeeping the module needs after the kernels ran (workspace views for hooks, CKY).

    The batch is processed as ``len(parts)`` independent chains (contiguous sentence ranges) on separate
    CUDA streams; each chain has its own workspace."""

    def __init__(self):
        self.parts = []          # [(b0, b1, ws, layout)]
        self.B = self.n = self.D = self.R = None
        self.consumed = False

    # first chain's workspace/layout (single-chain callers and tests)
    @property
    def ws(self):
        return self.parts[0][2]

    @property
    def layout(self):
        return self.parts[0][3]

    def _level_view(self, level, outside, what):
        L = _lib.lib()
        n, D = self.n, self.D
        Lc = n - level
        N = (n - level - 1) if outside else level
        out = []
        for b0, b1, ws, lay in self.parts:
            Bp = b1 - b0
            r0 = int(L.cliora_split_row_offset(Bp, n, level, 1 if outside else 0))
            rows = Bp * Lc * N
            if what == 'E':
                base = lay.Eout if outside else lay.Ein
                out.append(ws[base + r0: base + r0 + rows])
            else:
                base = {'Y': (lay.Yin, lay.Yout), 'Z': (lay.Zin, lay.Zout)}[what][1 if outside else 0]
                out.append(ws[base + r0 * D: base + (r0 + rows) * D].view(rows, D))
        return out[0] if len(out) == 1 else torch.cat(out, 0)

    def split_h(self, level, outside=False):
        """Pre-aggregation vectors of a level, [B*L*N, D] -- the ``h`` of inside_hook (diora.py:331)."""
        return self._level_view(level, outside, 'Y')

    def split_z(self, level, outside=False):
        """Hidden activations relu(W1 [l;r] + b1) of a level, [rows, D] (tf32-rounded part when the
        tensor-core path stores split pairs; its sign pattern is exact)."""
        return self._level_view(level, outside, 'Z')

    def split_s(self, level, outside=False):
        """Raw split scores of a level, [B,L,N,1] inside / [B,N,L,1] outside -- the ``s`` of the hooks."""
        s = self._level_view(level, outside, 'E')
        Lc = self.n - level
        return s.view(self.B, -1, Lc, 1) if outside else s.view(self.B, Lc, -1, 1)


def _weights_struct(cls, tensors, share):
    s = cls()
    keys = _SHARED_KEYS + ([] if share else _OUTSIDE_KEYS)
    for k, t in zip(keys, tensors):
        setattr(s, k, ptr(t))
    return s


_side_streams = {}


def _streams(dev, k):
    pool = _side_streams.setdefault(dev.index, [])
    while len(pool) < k:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:k]


def _ranges(B, chains):
    chains = max(1, min(int(chains), B))
    base, rem = divmod(B, chains)
    out, b0 = [], 0
    for i in range(chains):
        b1 = b0 + base + (1 if i < rem else 0)
        out.append((b0, b1))
        b0 = b1
    return out


def _off(t, nelem):
    """Device pointer of ``t`` advanced by nelem elements (None stays NULL)."""
    if t is None:
        return None
    return ptr(t) + nelem * t.element_size()


class ChartFunction(torch.autograd.Function):
    """(x, obj, weights) -> (inside_h, inside_s, outside_h, outside_s) with a hand-written backward.

    Sentences are independent, and at batch 32 every kernel of the 76-level chain is latency-bound; the batch
    is therefore cut into ``chains`` contiguous ranges that run as independent chains on side streams
    (fork/join around the calls; under CUDA-graph capture they become parallel branches)."""

    @staticmethod
    def forward(ctx, run, share, outside, chains, flags, x, obj, keep, *weights):
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
        dev = x.device
        ranges = _ranges(B, chains)
        with torch.cuda.device(dev):
            inside_h = torch.empty(B, C, D, device=dev, dtype=torch.float32)
            inside_s = torch.empty(B, C, 1, device=dev, dtype=torch.float32)
            if outside:
                outside_h = torch.empty(B, C, D, device=dev, dtype=torch.float32)
                outside_s = torch.empty(B, C, 1, device=dev, dtype=torch.float32)
            else:  # the reference leaves the outside chart at its zero fill (diora.py:19-22)
                outside_h = torch.zeros(B, C, D, device=dev, dtype=torch.float32)
                outside_s = torch.zeros(B, C, 1, device=dev, dtype=torch.float32)
            W = _weights_struct(Weights, weights, share)
            parts = []
            for b0, b1 in ranges:
                lay = _lib.layout(b1 - b0, n, D, R, share)
                parts.append((b0, b1, torch.empty(int(lay.ws_floats), device=dev, dtype=torch.float32), lay))
            cur = torch.cuda.current_stream(dev)
            streams = [cur] if len(parts) == 1 else _streams(dev, len(parts))
            for (b0, b1, ws, lay), st in zip(parts, streams):
                if st is not cur:
                    st.wait_stream(cur)
                dims = Dims(b1 - b0, n, D, R, 1 if share else 0, flags)
                with torch.cuda.stream(st):
                    h = st.cuda_stream
                    check(L.cliora_inside_fwd(ctypes.byref(dims), ctypes.byref(W), _off(x, b0 * n * D),
                                              _off(obj, b0 * R * D), _off(keep, b0 * C * R),
                                              _off(inside_h, b0 * C * D), _off(inside_s, b0 * C), ptr(ws), h),
                          'cliora_inside_fwd')
                    if outside:
                        check(L.cliora_outside_fwd(ctypes.byref(dims), ctypes.byref(W), _off(inside_h, b0 * C * D),
                                                   _off(inside_s, b0 * C), _off(outside_h, b0 * C * D),
                                                   _off(outside_s, b0 * C), ptr(ws), h), 'cliora_outside_fwd')
            for st in streams:
                if st is not cur:
                    cur.wait_stream(st)
        run.parts, run.B, run.n, run.D, run.R = parts, B, n, D, R
        ctx.run, ctx.share, ctx.outside, ctx.dims_t, ctx.flags = run, share, outside, (B, n, D, R), flags
        ctx.has_obj, ctx.has_keep = obj is not None, keep is not None
        saved = [x, inside_h, inside_s, outside_h, outside_s]
        if obj is not None:
            saved.append(obj)
        if keep is not None:
            saved.append(keep)
        ctx.n_ws = len(parts)
        ctx.save_for_backward(*saved, *[p[2] for p in parts], *weights)
        return inside_h, inside_s, outside_h, outside_s

    @staticmethod
    def backward(ctx, g_ih, g_is, g_oh, g_os):
        L = _lib.lib()
        if ctx.run.consumed:
            raise _lib.ClioraError('cliora_b200: the chart backward consumes its workspace; '
                                   'backward through the same forward twice is not supported')
        saved = list(ctx.saved_tensors)
        x, inside_h, inside_s, outside_h, outside_s = saved[:5]
        i = 5
        obj = keep = None
        if ctx.has_obj:
            obj = saved[i]; i += 1
        if ctx.has_keep:
            keep = saved[i]; i += 1
        wss = saved[i:i + ctx.n_ws]
        weights = saved[i + ctx.n_ws:]
        B, n, D, R = ctx.dims_t
        C = n * (n + 1) // 2
        share, outside = ctx.share, ctx.outside
        dev = x.device
        cont = lambda g: None if g is None else g.contiguous().float()
        g_ih, g_is, g_oh, g_os = cont(g_ih), cont(g_is), cont(g_oh), cont(g_os)
        if not outside:
            g_oh = g_os = None
        parts = ctx.run.parts
        with torch.cuda.device(dev):
            gx = torch.empty_like(x)
            gobj = torch.empty_like(obj) if obj is not None else None
            W = _weights_struct(Weights, weights, share)
            cur = torch.cuda.current_stream(dev)
            streams = [cur] if len(parts) == 1 else _streams(dev, len(parts))
            # scratch and per-chain weight grads are allocated on the current stream BEFORE the fork, so the
            # caching allocator ties their lifetime to `cur` (all side-stream work is joined back into it)
            all_bws = [torch.empty(int(p[3].bws_floats), device=dev, dtype=torch.float32) for p in parts]
            all_grads = [[torch.empty_like(w) for w in weights] for _ in parts]
            aux_streams = _streams(dev, 2 * len(parts))[len(parts):]
            for (b0, b1, _, lay), ws, st, bws, grads, aux in zip(parts, wss, streams, all_bws, all_grads, aux_streams):
                if st is not cur:
                    st.wait_stream(cur)
                dims = Dims(b1 - b0, n, D, R, 1 if share else 0, ctx.flags)
                with torch.cuda.stream(st):
                    h = st.cuda_stream
                    G = _weights_struct(WeightGrads, grads, share)
                    check(L.cliora_chart_bwd_begin(ctypes.byref(dims), _off(g_ih, b0 * C * D), _off(g_is, b0 * C),
                                                   _off(g_oh, b0 * C * D), _off(g_os, b0 * C), ptr(bws), h),
                          'cliora_chart_bwd_begin')
                    def outside_bwd(phase, handle):
                        check(L.cliora_outside_bwd(ctypes.byref(dims), ctypes.byref(W), _off(inside_h, b0 * C * D),
                                                   _off(inside_s, b0 * C), _off(outside_h, b0 * C * D),
                                                   _off(outside_s, b0 * C), ptr(ws), ptr(bws), ctypes.byref(G), phase,
                                                   handle), 'cliora_outside_bwd')

                    def inside_bwd(phase):
                        check(L.cliora_inside_bwd(ctypes.byref(dims), ctypes.byref(W), _off(x, b0 * n * D),
                                                  _off(obj, b0 * R * D), _off(keep, b0 * C * R),
                                                  _off(inside_h, b0 * C * D), _off(inside_s, b0 * C),
                                                  _off(outside_h, b0 * C * D), ptr(ws), ptr(bws), 1 if outside else 0,
                                                  _off(gx, b0 * n * D), _off(gobj, b0 * R * D), ctypes.byref(G), phase,
                                                  h), 'cliora_inside_bwd')

                    if outside:
                        # the outside pass's weight-gradient GEMMs (full-GPU tensor-core work) only need the outside
                        # level chain: run them on an auxiliary stream while the latency-bound inside level chain
                        # proceeds, and join before the inside weight phase accumulates into the same tensors
                        outside_bwd(1, h)
                        aux.wait_stream(st)
                        with torch.cuda.stream(aux):
                            outside_bwd(2, aux.cuda_stream)
                        inside_bwd(1)
                        st.wait_stream(aux)
                        inside_bwd(2)
                    else:
                        inside_bwd(3)
            for st in streams:
                if st is not cur:
                    cur.wait_stream(st)
            grads = all_grads[0]
            for other in all_grads[1:]:
                torch._foreach_add_(grads, other)
        ctx.run.consumed = True
        return (None, None, None, None, None, gx, gobj, None) + tuple(grads)
