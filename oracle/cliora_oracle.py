"""CPU oracle for the CLIORA chart hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``cliora_b200/`` imports it.

It restates, in plain torch-on-CPU (dtype-generic, so it can run in float64 as
a high-precision arbiter), the algorithm of the reference's chart path in the
reference's own *dense* formulation (cat -> W1 -> W2, bilinear as two matmuls,
gather by index tensors).  Gradients come from torch autograd over this
restatement.  Every function cites the reference file:line it follows
(paths relative to /root/reference).

Parity pinning: ``tests/golden/make_golden.py`` imports the unmodified
reference in the build container, runs it on seeded inputs and commits the
outputs (chart tensors, losses, gradients, CKY trees, index tensors) as
fixtures; ``tests/test_oracle_vs_golden.py`` checks this oracle against every
one of them, and ``tests/test_oracle_step_golden.py`` checks ``CpuClioraStep``
against three steps of the reference's own ``build_net`` + ``Trainer.step``
(``tests/golden/make_golden_step.py``).  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so those fixtures are the pin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

TINY = 1e-8  # cliora/net/utils.py:10


# --------------------------------------------------------------------------
# chart geometry (closed forms; the reference builds python lists)
# --------------------------------------------------------------------------

def num_cells(n: int) -> int:
    """cliora/net/diora.py:11  ncells = n(n+1)/2."""
    return n * (n + 1) // 2


def level_offsets(n: int) -> List[int]:
    """cliora/net/offset_cache.py:1-7.  offset[l] = ncells(n) - ncells(n-l)."""
    return [num_cells(n) - num_cells(n - l) for l in range(n)]


def inside_pairs(n: int, level: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """cliora/net/inside_index.py:131-197 (get_inside_components/get_inside_index).

    For every cell (level, p), p in [0, n-level), and split k in [0, level):
    left child = (k, p), right child = (level-1-k, p+k+1).  Flattened (p, k).
    """
    off = level_offsets(n)
    left, right = [], []
    for p in range(n - level):
        for k in range(level):
            left.append(off[k] + p)
            right.append(off[level - 1 - k] + p + k + 1)
    return torch.tensor(left, dtype=torch.int64), torch.tensor(right, dtype=torch.int64)


def outside_pairs(n: int, level: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """cliora/net/outside_index.py:39-62,93-127 (get_all_pairs/get_outside_index).

    Cell (level, p) has N = n-level-1 (parent, sibling) entries; the flattened
    order is (k, p) ("N-major").  For position p the first R = n-1-level-p
    entries have the sibling on the right: k-th sibling = (R-1-k, p+level+1),
    parent = (level+R-k, p); the remaining p entries have the sibling on the
    left: j-th sibling = (j, p-1-j), parent = (level+j+1, p-1-j).
    """
    off = level_offsets(n)
    L = n - level
    N = L - 1
    par = [[0] * L for _ in range(N)]
    sib = [[0] * L for _ in range(N)]
    for p in range(L):
        R = n - 1 - level - p
        for k in range(N):
            if k < R:
                s_lvl, s_pos = R - 1 - k, p + level + 1
                p_lvl, p_pos = level + R - k, p
            else:
                j = k - R
                s_lvl, s_pos = j, p - 1 - j
                p_lvl, p_pos = level + j + 1, p - 1 - j
            par[k][p] = off[p_lvl] + p_pos
            sib[k][p] = off[s_lvl] + s_pos
    flat = lambda rows: torch.tensor([v for r in rows for v in r], dtype=torch.int64)
    return flat(par), flat(sib)


# --------------------------------------------------------------------------
# small modules
# --------------------------------------------------------------------------

def unit(x: torch.Tensor) -> torch.Tensor:
    """cliora/net/utils.py:11-14  x / max(||x||_2, 1e-8)."""
    return x / x.norm(p=2, dim=-1, keepdim=True).clamp(min=TINY)


def normalize(x: torch.Tensor, mode: str) -> torch.Tensor:
    """cliora/net/utils.py:17-27."""
    return unit(x) if mode == 'unit' else x


def region_attention(q, obj, keep=None, p_drop=0.1):
    """cliora/net/cliora.py:28-42 (AttentionHead.forward).

    The reference forms the full [B,B,L,R] einsum and keeps the batch diagonal;
    only the diagonal is restated.  ``keep`` is an explicit {0,1} dropout mask
    [B,L,R] (None = eval mode / dropout off); kept entries are scaled 1/(1-p).
    """
    logits = torch.einsum('blx,brx->blr', q, obj)
    prob = torch.softmax(logits, dim=-1)
    if keep is not None:
        prob = prob * keep.to(prob.dtype) / (1.0 - p_drop)
    return torch.bmm(prob, obj)


def compose(P: Dict[str, torch.Tensor], prefix: str, a: torch.Tensor, b: torch.Tensor, pre=None):
    """cliora/net/diora.py:65-72 (ComposeMLP.forward): ReLU(W2 ReLU(W1 [a;b] + b1) + b2)."""
    x = torch.cat([a, b], dim=1)
    u1 = torch.addmm(P[prefix + '.h_fcs.0.bias'], x, P[prefix + '.h_fcs.0.weight'].t())
    u2 = torch.addmm(P[prefix + '.h_fcs.2.bias'], torch.relu(u1), P[prefix + '.h_fcs.2.weight'].t())
    if pre is not None:
        pre.append((u1.detach(), u2.detach()))
    return torch.relu(u2)


def bilinear(mat: torch.Tensor, a: torch.Tensor, b: torch.Tensor):
    """cliora/net/diora.py:89-97 (Bilinear.forward): a^T M b per row -> [M,1]."""
    return ((a @ mat) * b).sum(dim=1, keepdim=True)


# --------------------------------------------------------------------------
# the chart passes
# --------------------------------------------------------------------------

class ChartOut(object):
    """Mirrors cliora/net/diora.py:7-23 (Chart) minus the dead *_c tensors."""

    def __init__(self):
        self.inside_h = self.inside_s = self.outside_h = self.outside_s = None
        self.split_scores = {}   # level -> [B, L, N, 1] raw inside split scores (inside_hook's ``s``)
        self.split_h = {}        # level -> [B*L*N, D]  pre-aggregation vectors (inside_hook's ``h``)
        self.out_split_scores = {}
        self.pre_in = {}         # level -> (u1, u2): pre-activations of the two ReLUs (kink detection in tests)
        self.pre_out = {}


def leaf_transform(P, x, obj=None, keep=None, mode='unit'):
    """Text: cliora/net/diora.py:58-63,283-292.  VL: cliora/net/cliora.py:71-80,290-301."""
    w, b = P['inside_compose_func.leaf_fc.weight'], P['inside_compose_func.leaf_fc.bias']
    h = torch.tanh(x @ w.t() + b)
    if obj is None:
        return normalize(h, mode)
    h = normalize(h, mode)
    h = h + region_attention(h, obj, keep)
    return normalize(h, mode)


def inside_pass(P, leaf_h, obj=None, keep=None, mode='unit', out: Optional[ChartOut] = None):
    """cliora/net/diora.py:295-331 (+ :102-149 level functions); VL aggregate cliora/net/cliora.py:140-157.

    ``keep``: optional [B, ncells, R] dropout keep-mask indexed by chart cell.
    Returns inside_h [B,cells,D], inside_s [B,cells,1].
    """
    B, n, D = leaf_h.shape
    off = level_offsets(n)
    hs = [leaf_h]
    ss = [leaf_h.new_zeros(B, n, 1)]
    for level in range(1, n):
        L, N = n - level, level
        H = torch.cat(hs, 1)
        S = torch.cat(ss, 1)
        li, ri = (t.to(H.device) for t in inside_pairs(n, level))
        lh, rh = H.index_select(1, li).reshape(-1, D), H.index_select(1, ri).reshape(-1, D)
        ls, rs = S.index_select(1, li).reshape(-1, 1), S.index_select(1, ri).reshape(-1, 1)
        pre = [] if out is not None else None
        h = compose(P, 'inside_compose_func', lh, rh, pre)
        if out is not None:
            out.pre_in[level] = pre[0]
        s = (bilinear(P['inside_score_func.mat'], lh, rh) + ls + rs).view(B, L, N, 1)
        p = torch.softmax(s, dim=2)
        hbar = normalize((h.view(B, L, N, D) * p).sum(2), mode)
        sbar = (s * p).sum(2)
        if obj is not None:
            k = None if keep is None else keep[:, off[level]:off[level] + L]
            hbar = normalize(hbar + region_attention(hbar, obj, k), mode)
        if out is not None:
            out.split_scores[level] = s
            out.split_h[level] = h
        hs.append(hbar)
        ss.append(sbar)
    return torch.cat(hs, 1), torch.cat(ss, 1)


def outside_pass(P, inside_h, inside_s, mode='unit', out: Optional[ChartOut] = None):
    """cliora/net/diora.py:337-398 (+ :154-200); arguments are always (sibling_inside, parent_outside)."""
    B, C, D = inside_h.shape
    n = int((math.isqrt(8 * C + 1) - 1) // 2)
    off = level_offsets(n)
    root = normalize(P['root_vector_out_h'].view(1, 1, D).expand(B, 1, D), mode)
    # Levels are produced top-down; keep them in a dict and assemble at the end.
    lvl_h = {n - 1: root}
    lvl_s = {n - 1: inside_h.new_zeros(B, 1, 1)}

    def assemble(levels, width):
        parts = []
        for l in range(n):
            parts.append(levels[l] if l in levels else inside_h.new_zeros(B, n - l, width))
        return torch.cat(parts, 1)

    for level in range(n - 2, -1, -1):
        L = n - level
        OH, OS = assemble(lvl_h, D), assemble(lvl_s, 1)
        pi, si = (t.to(OH.device) for t in outside_pairs(n, level))
        ph, sh = OH.index_select(1, pi).reshape(-1, D), inside_h.index_select(1, si).reshape(-1, D)
        ps, ss = OS.index_select(1, pi).reshape(-1, 1), inside_s.index_select(1, si).reshape(-1, 1)
        pre = [] if out is not None else None
        h = compose(P, 'outside_compose_func', sh, ph, pre)
        if out is not None:
            out.pre_out[level] = pre[0]
        s = (bilinear(P['outside_score_func.mat'], sh, ph) + ss + ps).view(B, -1, L, 1)
        p = torch.softmax(s, dim=1)
        N = s.shape[1]
        lvl_h[level] = normalize((h.view(B, N, L, D) * p).sum(1), mode)
        lvl_s[level] = (s * p).sum(1)
        if out is not None:
            out.out_split_scores[level] = s
    return assemble(lvl_h, D), assemble(lvl_s, 1)


def chart_forward(P, x_span, obj_span=None, keep=None, outside=True, mode='unit') -> ChartOut:
    """cliora/net/diora.py:424-450 / cliora/net/cliora.py:438-455 (DioraBase.forward up to the einsums)."""
    out = ChartOut()
    n = x_span.shape[1]
    leaf_keep = None if keep is None else keep[:, :n]
    leaf = leaf_transform(P, x_span, obj_span, leaf_keep, mode)
    out.inside_h, out.inside_s = inside_pass(P, leaf, obj_span, keep, mode, out)
    if outside:
        out.outside_h, out.outside_s = outside_pass(P, out.inside_h, out.inside_s, mode, out)
    else:
        out.outside_h = torch.zeros_like(out.inside_h)
        out.outside_s = torch.zeros_like(out.inside_s)
    return out


# --------------------------------------------------------------------------
# span-region alignment and the three losses
# --------------------------------------------------------------------------

def all_atten_score(inside_h, outside_h, obj_span):
    """cliora/net/cliora.py:457  einsum('abx,cdx->acbd') -> [B_sent, B_img, cells, R]."""
    return torch.einsum('abx,cdx->acbd', inside_h + outside_h, obj_span)


def vg_atten_score(x_word, obj_word, training=True, all_atten=None, mode='unit'):
    """cliora/net/cliora.py:459-466."""
    if training:
        return torch.einsum('abx,cdx->acbd', x_word, obj_word)
    word = torch.einsum('abx,cdx->acbd', normalize(x_word, mode), obj_word)
    return all_atten[:, :, :x_word.shape[1]] + word


def atten_score(vg):
    """cliora/net/cliora.py:466  batch diagonal -> [B, n, R]."""
    return torch.diagonal(vg, 0, 0, 1).permute(2, 0, 1)


def contrastive_loss(all_atten, inside_s, outside_s, margin=0.2, alpha=1.0):
    """cliora/net/trainer.py:91-128 (ContrastiveLoss.forward)."""
    B = all_atten.shape[0]
    ins, outs = inside_s.squeeze(-1), outside_s.squeeze(-1)
    C = ins.shape[1]
    sc = all_atten.max(-1).values.permute(2, 0, 1)            # [cells, sent, img]
    diag = torch.diagonal(sc, 0, -1).unsqueeze(-1)             # [cells, B, 1]
    txt = (margin + sc - diag).clamp(min=TINY)
    img = (margin + sc - diag.transpose(1, 2)).clamp(min=TINY)
    eye = torch.eye(B, dtype=torch.bool, device=sc.device).unsqueeze(0)
    txt = txt.masked_fill(eye, 0).mean(2)
    img = img.masked_fill(eye, 0).mean(1)
    vl = (txt + img).t()                                       # [B, cells]
    marg = torch.exp(ins + outs - ins[:, [-1]])
    return (marg * vl)[:, :C // 2].sum(-1).mean() * alpha


def vg_loss(vg, alpha=1.0):
    """cliora/net/trainer.py:139-171 (VGLoss.forward, the live 'V1' branch)."""
    B, _, n, _ = vg.shape
    logits = vg.max(-1).values.sum(-1) / n
    return alpha * torch.nn.functional.cross_entropy(logits, torch.arange(B, device=logits.device))


def reconstruction_loss(emb_weight, mat, sentences, neg_samples, outside_h):
    """cliora/net/trainer.py:46-78 (ReconstructionSoftmaxLoss.forward)."""
    B, n = sentences.shape
    cell = outside_h[:, :n]                                    # [B,n,D]
    pos = emb_weight[sentences] @ mat.t()                      # [B,n,D]
    neg = emb_weight[neg_samples] @ mat.t()                    # [K,D]
    xp = (pos * cell).sum(-1, keepdim=True)
    xn = cell @ neg.t()
    score = torch.cat([xp, xn], 2).reshape(B * n, -1)
    tgt = torch.zeros(B * n, dtype=torch.int64, device=score.device)
    return torch.nn.functional.cross_entropy(score, tgt)


def embed(emb_weight, mat, mat1, sentences):
    """cliora/net/trainer.py:219-224 (Embed.forward)."""
    e = emb_weight[sentences.reshape(-1)]
    B, n = sentences.shape
    return (e @ mat.t()).view(B, n, -1), (e @ mat1.t()).view(B, n, -1)


def image_encoder(P, obj_feats):
    """cliora/net/utils.py:52-55 (ImageEncoder.forward)."""
    f = obj_feats.to(P['fc.weight'].dtype)
    return f @ P['fc.weight'].t() + P['fc.bias'], f @ P['fc_vis.weight'].t() + P['fc_vis.bias']


# --------------------------------------------------------------------------
# CKY
# --------------------------------------------------------------------------

def cky_backpointers(split_scores: Dict[int, torch.Tensor], B: int, n: int):
    """cliora/analysis/cky.py:31-99 with the hook of cliora/analysis/utils.py:78-95.

    ``split_scores[level]`` is the raw inside score [B, L, N, 1]; the hook
    subtracts the per-cell max before CKY sees it.  Leaves score 1.0
    (cky.py:23-24,40).  First-max tie-break (torch.argmax).  Returns
    (best [B, cells], backptr int32 [B, cells]) with backptr = split k, -1 at leaves.
    """
    off = level_offsets(n)
    C = num_cells(n)
    best = torch.ones(B, C, dtype=torch.float32)
    bp = torch.full((B, C), -1, dtype=torch.int32)
    for level in range(1, n):
        L, N = n - level, level
        s = split_scores[level].reshape(B, L, N).to(torch.float32)
        s = s - s.max(2, keepdim=True)[0]
        for p in range(L):
            cand = torch.stack([best[:, off[k] + p] + best[:, off[level - 1 - k] + p + k + 1] + s[:, p, k]
                                for k in range(N)], 1)
            am = cand.argmax(1)
            best[:, off[level] + p] = cand[torch.arange(B), am]
            bp[:, off[level] + p] = am.to(torch.int32)
    return best, bp


def tree_from_backpointers(bp_row: Sequence[int], n: int):
    """cliora/analysis/cky.py:101-109 (follow_backpointers): nested tuples of word positions."""
    off = level_offsets(n)

    def rec(level, pos):
        if level == 0:
            return pos
        k = int(bp_row[off[level] + pos])
        return (rec(k, pos), rec(level - 1 - k, pos + k + 1))

    return rec(n - 1, 0)


def cky_trees(split_scores, B, n):
    _, bp = cky_backpointers(split_scores, B, n)
    return [tree_from_backpointers(bp[b].tolist(), n) for b in range(B)]


# --------------------------------------------------------------------------
# a whole CLIORA / DIORA training step on CPU (used as bench.py's CPU baseline)
# --------------------------------------------------------------------------

def init_params(D: int, share=True, seed=0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference init: every Diora tensor ~ N(0,1) (cliora/net/diora.py:234-237)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    P = {
        'root_vector_out_h': rn(D),
        'inside_score_func.mat': rn(D, D),
        'inside_compose_func.leaf_fc.weight': rn(D, D),
        'inside_compose_func.leaf_fc.bias': rn(D),
        'inside_compose_func.h_fcs.0.weight': rn(D, 2 * D),
        'inside_compose_func.h_fcs.0.bias': rn(D),
        'inside_compose_func.h_fcs.2.weight': rn(D, D),
        'inside_compose_func.h_fcs.2.bias': rn(D),
    }
    if share:
        for k in list(P):
            if k.startswith('inside_'):
                P['outside_' + k[len('inside_'):]] = P[k]
    else:
        P['outside_score_func.mat'] = rn(D, D)
        for k in ('h_fcs.0.weight', 'h_fcs.0.bias', 'h_fcs.2.weight', 'h_fcs.2.bias'):
            P['outside_compose_func.' + k] = rn(*P['inside_compose_func.' + k].shape)
    return P


class CpuClioraStep(object):
    """One full CLIORA training step on CPU in the reference's own dense formulation
    (Net.forward -> losses -> backward -> clip 5.0 -> Adam; cliora/net/trainer.py:272-304,450-455,483-501).
    Used ONLY as bench.py's cpu_baseline / ``--impl reference`` arm and by tests as a checker."""

    def __init__(self, D=400, E=1024, V=8000, F=2048, k_neg=100, seed=1234, lr=2e-3, obj_feats=True,
                 alpha_vg=1.0, alpha_contr=1.0, margin=0.2, device='cpu', dtype=torch.float32):
        # dtype=torch.float64 makes this step the arbiter for gradient checks at deep charts (same values, drawn
        # in float32 and widened).  device != 'cpu' runs the same dense eager formulation on that device (the "stock PyTorch on the same
        # GPU" number of SURVEY.md section 8(d)); still a checker/baseline, never a product path
        g = torch.Generator().manual_seed(seed)
        self.obj_feats = obj_feats
        self.dtype = dtype
        self.P = {k: v.to(device, dtype).requires_grad_() for k, v in init_params(D, share=True, seed=seed).items()
                  if not k.startswith('outside_')}
        for k in list(self.P):
            if k.startswith('inside_'):
                self.P['outside_' + k[len('inside_'):]] = self.P[k]
        dev = lambda t: t.to(device, dtype)
        self.emb = dev(torch.randn(V, E, generator=g))                  # frozen with --obj_feats (trainer.py:538-541)
        self.mat = dev(torch.randn(D, E, generator=g)).requires_grad_()
        self.mat1 = dev(torch.randn(D, E, generator=g)).requires_grad_()
        self.recon_mat = dev(torch.randn(D, E, generator=g)).requires_grad_()
        self.enc = {'fc.weight': dev(0.02 * torch.randn(D, F, generator=g)).requires_grad_(),
                    'fc.bias': dev(0.02 * torch.randn(D, generator=g)).requires_grad_(),
                    'fc_vis.weight': dev(0.02 * torch.randn(D, F, generator=g)).requires_grad_(),
                    'fc_vis.bias': dev(0.02 * torch.randn(D, generator=g)).requires_grad_()}
        self.alpha_vg, self.alpha_contr, self.margin = alpha_vg, alpha_contr, margin
        uniq = {id(v): v for v in list(self.P.values()) + [self.mat, self.mat1, self.recon_mat] + list(self.enc.values())}
        self.params = list(uniq.values())
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=(0.9, 0.999), eps=1e-8)

    def loss(self, sentences, neg_samples, obj_feats=None, keep=None):
        x_span, x_word = embed(self.emb, self.mat, self.mat1, sentences)
        if self.obj_feats:
            obj_span, obj_word = image_encoder(self.enc, obj_feats.to(self.dtype))
            out = chart_forward(self.P, x_span, obj_span, keep)
            aas = all_atten_score(out.inside_h, out.outside_h, obj_span)
            vg = vg_atten_score(x_word, obj_word, training=True)
            losses = [reconstruction_loss(self.emb, self.recon_mat, sentences, neg_samples, out.outside_h),
                      vg_loss(vg, self.alpha_vg),
                      contrastive_loss(aas, out.inside_s, out.outside_s, self.margin, self.alpha_contr)]
        else:
            out = chart_forward(self.P, x_span)
            losses = [reconstruction_loss(self.emb, self.recon_mat, sentences, neg_samples, out.outside_h)]
        return sum(losses), losses

    def step(self, sentences, neg_samples, obj_feats=None, keep=None):
        self.opt.zero_grad()
        total, losses = self.loss(sentences, neg_samples, obj_feats, keep)
        total.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 5.0)
        self.opt.step()
        return total.item(), [l.item() for l in losses]
