"""Text-only DIORA-MLP chart model: drop-in for ``cliora.net.diora.DioraMLP``.

Same constructor, ``state_dict`` keys, attributes and hooks as the reference
(cliora/net/diora.py:205-471); the forward is four calls into libcliora_b200.so
instead of ~15 000 ATen ops.
"""
import torch
import torch.nn as nn

from .. import _lib
from .chart import Chart, ChartFunction, ChartRun
from .index import Index
from .utils import NormalizeFunc


class Bilinear(nn.Module):
    """Parameter holder for the bilinear split score l^T M r (cliora/net/diora.py:77-97)."""

    def __init__(self, size):
        super().__init__()
        self.size = size
        self.mat = nn.Parameter(torch.FloatTensor(size, size))


class ComposeMLP(nn.Module):
    """Parameter holder for the composition MLP (cliora/net/diora.py:26-72): leaf_fc, h_fcs.0, h_fcs.2."""

    def __init__(self, size, ninput=2, leaf=False):
        super().__init__()
        self.size, self.ninput = size, ninput
        if leaf:
            self.leaf_fc = nn.Linear(size, size)
        self.h_fcs = nn.Sequential(nn.Linear(2 * size, size), nn.ReLU(), nn.Linear(size, size), nn.ReLU())


class DioraBase(nn.Module):
    compose_cls = ComposeMLP
    visual = False

    def __init__(self, size, word_mat=None, cate_mat=None, outside=True, normalize='unit', compress=False,
                 share=True):
        super().__init__()
        if normalize not in ('unit', 'none'):      # scripts/train.py:341: choices=('none', 'unit')
            raise ValueError("normalize must be 'unit' or 'none'; got %r" % (normalize,))
        if compress:
            raise NotImplementedError('compress=True is unreachable in the reference (trainer.py:552)')
        self.size = size
        self.normalize = normalize
        self.share = share
        self.outside = outside
        self.inside_normalize_func = NormalizeFunc(normalize)
        self.outside_normalize_func = NormalizeFunc(normalize)
        self.compress = compress
        self.ninput = 2
        self.index = None
        self.charts = None
        self._run = None
        self._pending = None
        self._keep_override = None
        self.chains = None     # concurrent sentence sub-batches (None: pick from the batch size)
        # 'fp32': tensor-core GEMMs are fp32-accurate (3xTF32, default, <= 1e-4 vs the reference);
        # 'tf32': single TF32 pass, stated tolerance 1e-2 of max, trees not guaranteed identical
        # 'bf16': bf16 operands in the compose GEMMs of the fused level kernels (fp32 accumulate), single-pass TF32
        #         elsewhere; stated tolerance 3e-2 of max on vectors, 1e-2 on scores; trees not guaranteed identical
        self.precision = 'fp32'
        # fused level kernels (one launch per level forward, one backward) or the unfused per-level chain: 'auto'
        # fuses whenever the shape is supported (measured on B200, n=20: 6336 vs 5826 sent/s at batch 32, 8568 vs 7955
        # at 64, 9445 vs 8780 at 128 -- levels that overflow one wave run with wide column slices);
        # True / False force one path; results are the same
        self.fused = 'auto'
        self.init_parameters()
        self.reset_parameters()
        self.reset()

    # ---- parameters: same modules, same registration order, same init as the reference ----
    def init_parameters(self):
        self.inside_score_func = Bilinear(self.size)
        self.inside_compose_func = self.compose_cls(self.size, leaf=True)
        if self.share:
            self.outside_score_func = self.inside_score_func
            self.outside_compose_func = self.inside_compose_func
        else:
            self.outside_score_func = Bilinear(self.size)
            self.outside_compose_func = self.compose_cls(self.size)
        self.root_vector_out_h = nn.Parameter(torch.FloatTensor(self.size))
        self.root_vector_out_c = None

    def reset_parameters(self):
        for p in self.parameters():   # N(0,1) for every tensor, diora.py:234-237
            if p.requires_grad:
                p.data.normal_()

    # ---- reference attribute surface ----
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def is_cuda(self):
        d = self.device
        return d.index is not None and d.index >= 0

    inside_h = property(lambda self: self.chart.inside_h)
    inside_c = property(lambda self: self.chart.inside_c)
    inside_s = property(lambda self: self.chart.inside_s)
    outside_h = property(lambda self: self.chart.outside_h)
    outside_c = property(lambda self: self.chart.outside_c)
    outside_s = property(lambda self: self.chart.outside_s)

    def cuda(self, device=None):
        super().cuda(device)
        if self.index is not None:
            self.index.cuda = True
        return self   # the reference returns None here (diora.py:272-275); returning self is a superset

    def get(self, chart, level):
        off = self.index.get_offset(self.length)[level]
        return chart[:, off:off + self.length - level]

    def get_chart_wrapper(self):
        return self

    def inside_hook(self, level, h, c, s):
        pass

    def outside_hook(self, level, h, c, s):
        pass

    def init_with_batch(self, h, c):
        """Reference signature (diora.py:401-410).  The chart tensors were produced by the kernels
        before this is called; ``h`` is the leaf slice of inside_h.  parse.py monkey-patches this on the
        instance to add ``saved_scalars`` (analysis/utils.py:67-75) and still works."""
        self.batch_size, self.length = h.shape[0], h.shape[1]
        self.chart = Chart(*self._pending)

    def reset(self):
        self.batch_size = None
        self.length = None
        self.chart = None
        self.atten_score = None
        self.all_atten_score = None
        self.vg_atten_score = None

    # ---- kernel bridge ----
    def _weight_list(self):
        ic, isf = self.inside_compose_func, self.inside_score_func
        w = [ic.leaf_fc.weight, ic.leaf_fc.bias, ic.h_fcs[0].weight, ic.h_fcs[0].bias, ic.h_fcs[2].weight,
             ic.h_fcs[2].bias, isf.mat, self.root_vector_out_h]
        if not self.share:
            oc, osf = self.outside_compose_func, self.outside_score_func
            w += [oc.h_fcs[0].weight, oc.h_fcs[0].bias, oc.h_fcs[2].weight, oc.h_fcs[2].bias, osf.mat]
        return w

    def _hook_overridden(self, name):
        fn = getattr(self, name)
        return getattr(fn, '__func__', fn) is not getattr(DioraBase, name)

    def set_dropout_mask(self, keep):
        """Testing hook: use this keep-mask [B, cells, R] for the next forward instead of drawing one."""
        self._keep_override = keep

    def run_chart(self, x_span, obj_embed_span=None):
        if self.index is None:
            self.index = Index(cuda=self.is_cuda)
        self.reset()
        B, n, _ = x_span.shape
        obj = obj_embed_span if self.visual else None
        keep = None
        if obj is not None and self.training:
            keep = self._keep_override
            p_drop = self.atten_head.dropout.p
            if p_drop not in (0.0, 0.1):
                raise NotImplementedError('the attention kernels hard-wire Dropout(0.1) (cliora.py:32); p=%r' % p_drop)
            if keep is None and p_drop > 0:   # one draw per forward instead of one per level
                keep = torch.rand(B, n * (n + 1) // 2, obj.shape[1], device=x_span.device) >= p_drop
        self._keep_override = None
        run = ChartRun()
        # measured on B200 (n=20, D=400): 2 chains are best at batch 32 (5034 vs 4910 sent/s with 4), 4 at batch 128
        chains = self.chains if self.chains is not None else max(1, min(2 if B < 64 else 4, B // 8))
        if self.precision not in ('fp32', 'tf32', 'bf16'):
            raise ValueError("precision must be 'fp32', 'tf32' or 'bf16'")
        fused = True if self.fused == 'auto' else bool(self.fused)
        flags = {'fp32': 0, 'tf32': 2, 'bf16': 8}[self.precision] | (0 if fused else 4) | ((min(chains, 15) & 15) << 8)
        if self.normalize == 'none':
            flags |= 16        # CLIORA_FLAG_NO_NORMALIZE
        outs = ChartFunction.apply(run, bool(self.share), bool(self.outside), chains, flags, x_span, obj, keep,
                                   *self._weight_list())
        self._run = run
        self._pending = outs
        self.init_with_batch(outs[0][:, :n], None)   # looked up dynamically: may be monkey-patched
        self._pending = None
        if self._hook_overridden('inside_hook'):
            for level in range(1, n):
                h = run.split_h(level)
                self.inside_hook(level, h, torch.zeros_like(h), run.split_s(level))
        if self.outside and self._hook_overridden('outside_hook'):
            for level in range(n - 2, -1, -1):
                h = run.split_h(level, outside=True)
                self.outside_hook(level, h, torch.zeros_like(h), run.split_s(level, outside=True))

    def forward(self, x_span, x_word=None, obj_embed_span=None, obj_embed_word=None):
        self.run_chart(x_span)
        return None


class DioraMLP(DioraBase):
    """``cliora.net.diora.DioraMLP`` (diora.py:453-471)."""
    pass
