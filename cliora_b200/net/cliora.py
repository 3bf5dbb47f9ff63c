"""Vision-language chart model: drop-in for ``cliora.net.cliora.DioraMLP`` (cliora/net/cliora.py).

Adds the per-level span->region attention (fused into the cell kernels) and the span-region
alignment scores.  ``all_atten_score`` / ``vg_atten_score`` / ``atten_score`` are materialised lazily,
only when a caller reads them (the fused losses in cliora_b200.net.trainer never do).
"""
import ctypes

import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, ptr
from .diora import DioraBase


class AttentionHead(nn.Module):
    """Parameter-free holder kept for module-tree parity (cliora.py:28-42); the attention itself runs
    inside cell_aggregate_kernel, diagonal (own image) only."""

    def __init__(self, q_dim, k_dim, v_dim, h_dim):
        super().__init__()
        self.h_dim = h_dim
        self.dropout = nn.Dropout(0.1)


class VLComposeMLP(nn.Module):
    """Parameter holder, same keys as cliora.py:46-86."""

    def __init__(self, size, ninput=2, leaf=False):
        super().__init__()
        self.size, self.ninput = size, ninput
        if leaf:
            self.leaf_fc = nn.Linear(size, size)
        self.h_fcs = nn.Sequential(nn.Linear(2 * size, size), nn.ReLU(), nn.Linear(size, size), nn.ReLU())


class AttenScores(torch.autograd.Function):
    """scores[a,c,cell,r] = h[a,cell] . obj[c,r]  (cliora.py:457/459), materialised."""

    @staticmethod
    def forward(ctx, h, obj):
        h, obj = h.contiguous().float(), obj.contiguous().float()
        B, ncell, D = h.shape
        R = obj.shape[1]
        out = torch.empty(B, B, ncell, R, device=h.device, dtype=torch.float32)
        with torch.cuda.device(h.device):
            check(_lib.lib().cliora_atten_scores(B, ncell, D, R, ptr(h), ncell, ptr(obj), ptr(out), _lib.stream()),
                  'cliora_atten_scores')
        ctx.save_for_backward(h, obj)
        return out

    @staticmethod
    def backward(ctx, g):
        # Compatibility path (only reached when a caller differentiates through the materialised
        # 4-D tensor, e.g. the reference's own ContrastiveLoss); the product path is AttenMax.
        h, obj = ctx.saved_tensors
        return torch.einsum('acbd,cdx->abx', g, obj), torch.einsum('acbd,abx->cdx', g, h)


class AttenMax(torch.autograd.Function):
    """smax[a,c,cell] = max_r h[a,cell] . obj[c,r] without materialising the scores."""

    @staticmethod
    def forward(ctx, h, obj):
        h, obj = h.contiguous().float(), obj.contiguous().float()
        B, ncell, D = h.shape
        R = obj.shape[1]
        smax = torch.empty(B, B, ncell, device=h.device, dtype=torch.float32)
        amax = torch.empty(B, B, ncell, device=h.device, dtype=torch.int32)
        L = _lib.lib()
        with torch.cuda.device(h.device):
            if D >= 32 and D % 4 == 0 and 2.0 * B * ncell * B * R * D >= 2e8:
                # tcgen05 3xTF32 GEMM with the max-over-regions epilogue (operands re-written as split pairs)
                hp = torch.empty(2, B * ncell, D, device=h.device, dtype=torch.float32)
                op = torch.empty(2, B * R, D, device=h.device, dtype=torch.float32)
                check(L.cliora_split_tf32(ptr(h), h.numel(), ptr(hp), _lib.stream()), 'cliora_split_tf32')
                check(L.cliora_split_tf32(ptr(obj), obj.numel(), ptr(op), _lib.stream()), 'cliora_split_tf32')
                check(L.cliora_tc_atten_max_fwd(B, ncell, D, R, ptr(hp), ptr(op), ptr(smax), ptr(amax), _lib.stream()),
                      'cliora_tc_atten_max_fwd')
            else:
                check(L.cliora_atten_max_fwd(B, ncell, D, R, ptr(h), ncell, ptr(obj), ptr(smax), ptr(amax),
                                             _lib.stream()), 'cliora_atten_max_fwd')
        ctx.save_for_backward(h, obj, amax)
        ctx.mark_non_differentiable(amax)
        return smax, amax

    @staticmethod
    def backward(ctx, g, _):
        h, obj, amax = ctx.saved_tensors
        B, ncell, D = h.shape
        R = obj.shape[1]
        g = g.contiguous().float()
        gh, gobj = torch.zeros_like(h), torch.zeros_like(obj)
        with torch.cuda.device(h.device):
            check(_lib.lib().cliora_atten_max_bwd(B, ncell, D, R, ptr(h), ncell, ptr(obj), ptr(g), ptr(amax),
                                                  ptr(gh), ncell, ptr(gobj), _lib.stream()), 'cliora_atten_max_bwd')
        return gh, gobj


class DioraMLP(DioraBase):
    """``cliora.net.cliora.DioraMLP``: ctor has no word_mat/cate_mat (cliora.py:216)."""
    compose_cls = VLComposeMLP
    visual = True

    def __init__(self, size, outside=True, normalize='unit', compress=False, share=True):
        super().__init__(size, outside=outside, normalize=normalize, compress=compress, share=share)

    def init_parameters(self):
        self.atten_head = AttentionHead(self.size, self.size, self.size, self.size)
        super().init_parameters()

    # lazily materialised alignment tensors -------------------------------------------------
    def reset(self):
        super().reset()
        self._lazy = {}
        self._x_word = self._obj_span = self._obj_word = None
        self.vg_atten_score_word = None

    def _get_lazy(self, name, fn):
        v = self.__dict__.get('_set_' + name)
        if v is not None:
            return v
        if self._obj_span is None:
            return None
        if name not in self._lazy:
            self._lazy[name] = fn()
        return self._lazy[name]

    def span_vectors(self):
        """inside_h + outside_h, the span representation scored against regions (cliora.py:457)."""
        return self._get_lazy('hsum', lambda: self.chart.inside_h + self.chart.outside_h)

    @property
    def all_atten_score(self):
        return self._get_lazy('all', lambda: AttenScores.apply(self.span_vectors(), self._obj_span))

    @all_atten_score.setter
    def all_atten_score(self, v):
        self.__dict__['_set_all'] = v

    def _vg(self):
        if self.training:   # cliora.py:459-461
            return AttenScores.apply(self._x_word, self._obj_word)
        n = self._x_word.shape[1]
        word = AttenScores.apply(self.inside_normalize_func(self._x_word), self._obj_word)
        return self.all_atten_score[:, :, :n] + word   # cliora.py:463-464

    @property
    def vg_atten_score(self):
        return self._get_lazy('vg', self._vg)

    @vg_atten_score.setter
    def vg_atten_score(self, v):
        self.__dict__['_set_vg'] = v

    @property
    def atten_score(self):
        return self._get_lazy('att', lambda: torch.diagonal(self.vg_atten_score, 0, 0, 1).permute(2, 0, 1))

    @atten_score.setter
    def atten_score(self, v):
        self.__dict__['_set_att'] = v

    # fused product path: what the losses in cliora_b200.net.trainer consume ---------------------
    def span_region_max(self, ncell=None):
        """max_r all_atten_score[..., r] for the first ``ncell`` cells -> ([B,B,ncell], argmax)."""
        h = self.span_vectors()
        if ncell is not None:
            h = h[:, :ncell]
        return AttenMax.apply(h, self._obj_span)

    def word_region_max(self):
        """max_r vg_atten_score_word[..., r] in training mode -> [B,B,n] (cliora.py:459, trainer.py:144)."""
        return AttenMax.apply(self._x_word, self._obj_word)

    def forward(self, x_span, x_word=None, obj_embed_span=None, obj_embed_word=None):
        if obj_embed_span is None:
            raise ValueError('cliora.DioraMLP needs obj_embed_span (use cliora_b200.net.diora.DioraMLP for text-only)')
        self.run_chart(x_span, obj_embed_span)
        self._x_word, self._obj_span, self._obj_word = x_word, obj_embed_span, obj_embed_word
        return None
