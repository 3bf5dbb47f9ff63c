"""Kernel-math emulator.  TEST INFRASTRUCTURE ONLY (same rules as cliora_oracle.py).

The CUDA kernels do not run the reference's dense formulation.  They run a
*factored* one (per-cell projections, per-split W2 GEMM) with a hand-derived
backward (no autograd).  This file restates exactly that algorithm -- same
buffers, same per-level phases, same closed-form index arithmetic -- in plain
torch on CPU so the derivation can be checked against autograd over the dense
oracle without a GPU.  Each phase names the CUDA kernel that implements it
(cliora_b200/csrc/).

Buffers (B sentences, n words, C=n(n+1)/2 cells, D hidden, R regions):
  inside_h [B,C,D] inside_s [B,C]   outside_h/outside_s likewise
  Pin  [B,C,PI*D]  per-inside-cell projections   (Al = h W1l^T, Ar = h W1r^T, V = h Wb^T [, Al' = h W1l'^T if !share])
  Pout [B,C,2D]    per-outside-cell projections  (Ar = h W1r'^T, V = h Wb'^T)
  Z/Y  [rows,D]    per-split hidden / output of the compose MLP, E/Pr [rows] split score / softmax prob
  per cell: q (pre-attention unit vector), nrm, nrm2, att [R]
"""
from __future__ import annotations

import torch

from .cliora_oracle import TINY, level_offsets, num_cells


def in_rows(n, level):
    """(left, right, cell) chart indices of every inside split row of a level, order (p,k)."""
    off = level_offsets(n)
    L, N = n - level, level
    p = torch.arange(L).view(L, 1).expand(L, N)
    k = torch.arange(N).view(1, N).expand(L, N)
    offt = torch.tensor(off)
    left = offt[k] + p
    right = offt[level - 1 - k] + p + k + 1
    cell = off[level] + p
    return left.reshape(-1), right.reshape(-1), cell.reshape(-1)


def out_rows(n, level):
    """(parent, sibling, cell) chart indices of every outside split row of a level, order (k,p)."""
    off = level_offsets(n)
    offt = torch.tensor(off)
    L = n - level
    N = L - 1
    k = torch.arange(N).view(N, 1).expand(N, L)
    p = torch.arange(L).view(1, L).expand(N, L)
    R = n - 1 - level - p
    right_side = k < R
    j = k - R
    s_lvl = torch.where(right_side, R - 1 - k, j)
    s_pos = torch.where(right_side, p + level + 1, p - 1 - j)
    p_lvl = torch.where(right_side, level + R - k, level + j + 1)
    p_pos = torch.where(right_side, p, p - 1 - j)
    par = offt[p_lvl.clamp(0, n - 1)] + p_pos
    sib = offt[s_lvl.clamp(0, n - 1)] + s_pos
    cell = off[level] + p
    return par.reshape(-1), sib.reshape(-1), cell.reshape(-1)


class Weights(object):
    """Splits the reference state_dict into the matrices the kernels consume."""

    def __init__(self, P, share=True):
        D = P['root_vector_out_h'].shape[0]
        self.D, self.share = D, share
        self.W_leaf, self.b_leaf = P['inside_compose_func.leaf_fc.weight'], P['inside_compose_func.leaf_fc.bias']
        W1 = P['inside_compose_func.h_fcs.0.weight']
        self.W1l, self.W1r, self.b1 = W1[:, :D], W1[:, D:], P['inside_compose_func.h_fcs.0.bias']
        self.W2, self.b2 = P['inside_compose_func.h_fcs.2.weight'], P['inside_compose_func.h_fcs.2.bias']
        self.Wb = P['inside_score_func.mat']
        self.root = P['root_vector_out_h']
        if share:
            self.oW1l, self.oW1r, self.ob1, self.oW2, self.ob2, self.oWb = \
                self.W1l, self.W1r, self.b1, self.W2, self.b2, self.Wb
        else:
            oW1 = P['outside_compose_func.h_fcs.0.weight']
            self.oW1l, self.oW1r, self.ob1 = oW1[:, :D], oW1[:, D:], P['outside_compose_func.h_fcs.0.bias']
            self.oW2, self.ob2 = P['outside_compose_func.h_fcs.2.weight'], P['outside_compose_func.h_fcs.2.bias']
            self.oWb = P['outside_score_func.mat']
        # [PI*D, D]: rows Al | Ar | V | (Al' when !share)
        parts = [self.W1l, self.W1r, self.Wb] + ([] if share else [self.oW1l])
        self.Wcat_in = torch.cat(parts, 0)
        self.Wcat_out = torch.cat([self.oW1r, self.oWb], 0)
        self.PI = len(parts)


def _finalize(a, obj, keep):
    """kernel: cell_finalize.  a [B,L,D] -> h, plus saved (q, nrm, nrm2, att)."""
    nrm = a.norm(dim=-1).clamp(min=TINY)
    q = a / nrm.unsqueeze(-1)
    if obj is None:
        return q, q, nrm, None, None
    att = torch.softmax(torch.einsum('bld,brd->blr', q, obj), -1)
    patt = att if keep is None else att * keep.to(att.dtype) / 0.9
    a2 = q + torch.bmm(patt, obj)
    nrm2 = a2.norm(dim=-1).clamp(min=TINY)
    return a2 / nrm2.unsqueeze(-1), q, nrm, nrm2, att


def _unit_bwd(g, h, nrm):
    """d/da of a/clamp(||a||,eps).  Clamped branch: the clamp passes no gradient -> g/eps."""
    live = (nrm > TINY).unsqueeze(-1)
    proj = g - h * (h * g).sum(-1, keepdim=True)
    return torch.where(live, proj, g) / nrm.unsqueeze(-1)


def _finalize_bwd(gh, h, q, nrm, nrm2, att, obj, keep, g_obj):
    """kernel: cell_finalize_bwd.  Returns grad wrt the pre-normalisation vector a; accumulates g_obj."""
    if obj is None:
        return _unit_bwd(gh, h, nrm)
    ga2 = _unit_bwd(gh, h, nrm2)
    scale = 1.0 if keep is None else keep.to(att.dtype) / 0.9
    patt = att * scale
    g_patt = torch.einsum('bld,brd->blr', ga2, obj)
    g_obj += torch.einsum('blr,bld->brd', patt, ga2)
    g_att = g_patt * scale
    g_logit = att * (g_att - (att * g_att).sum(-1, keepdim=True))
    gq = ga2 + torch.bmm(g_logit, obj)
    g_obj += torch.einsum('blr,bld->brd', g_logit, q)
    return _unit_bwd(gq, q, nrm)


class Saved(object):
    pass


def forward(P, x, obj=None, keep=None, outside=True, share=True):
    W = Weights(P, share)
    B, n, D = x.shape
    C = num_cells(n)
    off = level_offsets(n)
    sv = Saved()
    sv.W, sv.B, sv.n, sv.D, sv.C, sv.x, sv.obj, sv.keep, sv.outside = W, B, n, D, C, x, obj, keep, outside
    f = lambda *s: x.new_zeros(*s)
    ih, is_, oh, os_ = f(B, C, D), f(B, C), f(B, C, D), f(B, C)
    q_in, nrm_in = f(B, C, D), f(B, C)
    nrm2_in = f(B, C) if obj is not None else None
    att_in = f(B, C, obj.shape[1]) if obj is not None else None
    nrm_out = f(B, C)
    Pin, Pout = f(B, C, W.PI * D), f(B, C, 2 * D)

    def cells(level):
        return slice(off[level], off[level] + n - level)

    def fin_in(level, a):
        c = cells(level)
        k = None if keep is None else keep[:, c]
        h, q, nrm, nrm2, att = _finalize(a, obj, k)
        ih[:, c], q_in[:, c], nrm_in[:, c] = h, q, nrm
        if obj is not None:
            nrm2_in[:, c], att_in[:, c] = nrm2, att
        Pin[:, c] = h @ W.Wcat_in.t()                       # kernel: gemm (project)

    # leaf  (kernel: gemm + tanh epilogue, then cell_finalize)
    t = torch.tanh(x @ W.W_leaf.t() + W.b_leaf)
    fin_in(0, t)

    sv.in_lv = {}
    for level in range(1, n):
        L, N = n - level, level
        li, ri, _ = in_rows(n, level)
        # kernel: split_build  (z and score)
        Z = torch.relu(Pin[:, li, 0:D] + Pin[:, ri, D:2 * D] + W.b1)                  # [B, L*N, D]
        E = (ih[:, li] * Pin[:, ri, 2 * D:3 * D]).sum(-1) + is_[:, li] + is_[:, ri]   # [B, L*N]
        # kernel: gemm (W2, bias+relu epilogue)
        Y = torch.relu(Z @ W.W2.t() + W.b2)
        # kernel: cell_aggregate (softmax over splits, weighted sum) + cell_finalize
        Pr = torch.softmax(E.view(B, L, N), -1)
        a = (Y.view(B, L, N, D) * Pr.unsqueeze(-1)).sum(2)
        is_[:, cells(level)] = (E.view(B, L, N) * Pr).sum(-1)
        fin_in(level, a)
        sv.in_lv[level] = (Z, Y, E, Pr.reshape(B, L * N))

    sv.out_lv = {}
    if outside:
        r = W.root.view(1, 1, D).expand(B, 1, D)
        nr = r.norm(dim=-1).clamp(min=TINY)
        oh[:, C - 1:C] = r / nr.unsqueeze(-1)
        nrm_out[:, C - 1:C] = nr
        if n > 1:
            Pout[:, C - 1:C] = oh[:, C - 1:C] @ W.Wcat_out.t()
        iAl = 0 if share else 3 * D
        for level in range(n - 2, -1, -1):
            L, N = n - level, n - level - 1
            pi, si, _ = out_rows(n, level)
            Z = torch.relu(Pin[:, si, iAl:iAl + D] + Pout[:, pi, 0:D] + W.ob1)        # [B, N*L, D]
            E = (ih[:, si] * Pout[:, pi, D:2 * D]).sum(-1) + is_[:, si] + os_[:, pi]
            Y = torch.relu(Z @ W.oW2.t() + W.ob2)
            Pr = torch.softmax(E.view(B, N, L), 1)
            a = (Y.view(B, N, L, D) * Pr.unsqueeze(-1)).sum(1)
            c = cells(level)
            os_[:, c] = (E.view(B, N, L) * Pr).sum(1)
            nrm = a.norm(dim=-1).clamp(min=TINY)
            oh[:, c], nrm_out[:, c] = a / nrm.unsqueeze(-1), nrm
            if level > 0:
                Pout[:, c] = oh[:, c] @ W.Wcat_out.t()
            sv.out_lv[level] = (Z, Y, E, Pr.reshape(B, N * L))

    sv.inside_h, sv.inside_s, sv.outside_h, sv.outside_s = ih, is_, oh, os_
    sv.q_in, sv.nrm_in, sv.nrm2_in, sv.att_in, sv.nrm_out, sv.Pin, sv.Pout = \
        q_in, nrm_in, nrm2_in, att_in, nrm_out, Pin, Pout
    return sv


def _split_bwd(ga, gsbar, qa, nrm, sbar, Z, Y, E, Pr, W2, cell_of_row):
    """kernels: split_bwd (per row) then gemm (gy W2, mask z>0).

    ga [B,L,D] grad wrt the aggregated (pre-normalisation) vector, gsbar [B,L].
    cell_of_row maps each split row to its local cell index in [0,L).
    """
    a_dot_ga = (qa * ga).sum(-1) * nrm                         # a = q * nrm
    common = a_dot_ga + sbar * gsbar                           # sum_m p_m gp_m
    ga_r, gs_r, cm_r = ga[:, cell_of_row], gsbar[:, cell_of_row], common[:, cell_of_row]
    gp = (Y * ga_r).sum(-1) + E * gs_r
    ge = Pr * (gs_r + gp - cm_r)
    GY = Pr.unsqueeze(-1) * ga_r * (Y > 0)
    GZ = (GY @ W2) * (Z > 0)
    return GY, GZ, ge


def backward(sv, g_inside_h, g_inside_s, g_outside_h, g_outside_s):
    W, B, n, D, C = sv.W, sv.B, sv.n, sv.D, sv.C
    off = level_offsets(n)
    x, obj, keep = sv.x, sv.obj, sv.keep
    ih, is_, oh, os_ = sv.inside_h, sv.inside_s, sv.outside_h, sv.outside_s
    Pin, Pout = sv.Pin, sv.Pout
    f = lambda *s: x.new_zeros(*s)
    Gh_in, Gs_in = g_inside_h.clone(), g_inside_s.reshape(B, C).clone()
    Gh_out, Gs_out = g_outside_h.clone(), g_outside_s.reshape(B, C).clone()
    GP_in, GP_out = f(B, C, W.PI * D), f(B, C, 2 * D)
    g_obj = torch.zeros_like(obj) if obj is not None else None
    dW2, db2, db1 = f(D, D), f(D), f(D)
    doW2, dob2, dob1 = f(D, D), f(D), f(D)
    g_root = f(D)
    iAl = 0 if W.share else 3 * D

    def cells(level):
        return slice(off[level], off[level] + n - level)

    def scatter(dst, idx, src):
        dst.index_add_(1, idx, src)                            # kernel: split_scatter (red.global.add)

    if sv.outside:
        for level in range(0, n - 1):
            L, N = n - level, n - level - 1
            c = cells(level)
            gh = Gh_out[:, c] + (GP_out[:, c] @ W.Wcat_out if level > 0 else 0)   # kernel: gemm (cell grad)
            ga = _unit_bwd(gh, oh[:, c], sv.nrm_out[:, c])
            Z, Y, E, Pr = sv.out_lv[level]
            pi, si, ci = out_rows(n, level)
            GY, GZ, ge = _split_bwd(ga, Gs_out[:, c], oh[:, c], sv.nrm_out[:, c], os_[:, c],
                                    Z, Y, E, Pr, W.oW2, ci - off[level])
            doW2 += torch.einsum('brd,bre->de', GY, Z)
            dob2 += GY.sum((0, 1))
            dob1 += GZ.sum((0, 1))
            scatter(GP_in[:, :, iAl:iAl + D], si, GZ)
            scatter(GP_out[:, :, 0:D], pi, GZ)
            scatter(Gh_in, si, ge.unsqueeze(-1) * Pout[:, pi, D:2 * D])
            scatter(GP_out[:, :, D:2 * D], pi, ge.unsqueeze(-1) * ih[:, si])
            scatter(Gs_in, si, ge)
            scatter(Gs_out, pi, ge)
        # root of the outside chart = unit(root_vector)
        gh = Gh_out[:, C - 1:C] + (GP_out[:, C - 1:C] @ W.Wcat_out if n > 1 else 0)
        g_root = _unit_bwd(gh, oh[:, C - 1:C], sv.nrm_out[:, C - 1:C]).sum((0, 1))

    for level in range(n - 1, 0, -1):
        L, N = n - level, level
        c = cells(level)
        gh = Gh_in[:, c] + GP_in[:, c] @ W.Wcat_in
        k = None if keep is None else keep[:, c]
        ga = _finalize_bwd(gh, ih[:, c], sv.q_in[:, c], sv.nrm_in[:, c],
                           None if obj is None else sv.nrm2_in[:, c],
                           None if obj is None else sv.att_in[:, c], obj, k, g_obj)
        Z, Y, E, Pr = sv.in_lv[level]
        li, ri, ci = in_rows(n, level)
        GY, GZ, ge = _split_bwd(ga, Gs_in[:, c], sv.q_in[:, c], sv.nrm_in[:, c], is_[:, c],
                                Z, Y, E, Pr, W.W2, ci - off[level])
        dW2 += torch.einsum('brd,bre->de', GY, Z)
        db2 += GY.sum((0, 1))
        db1 += GZ.sum((0, 1))
        scatter(GP_in[:, :, 0:D], li, GZ)
        scatter(GP_in[:, :, D:2 * D], ri, GZ)
        scatter(Gh_in, li, ge.unsqueeze(-1) * Pin[:, ri, 2 * D:3 * D])
        scatter(GP_in[:, :, 2 * D:3 * D], ri, ge.unsqueeze(-1) * ih[:, li])
        scatter(Gs_in, li, ge)
        scatter(Gs_in, ri, ge)

    # leaves
    c = cells(0)
    gh = Gh_in[:, c] + GP_in[:, c] @ W.Wcat_in
    k = None if keep is None else keep[:, c]
    gt = _finalize_bwd(gh, ih[:, c], sv.q_in[:, c], sv.nrm_in[:, c],
                       None if obj is None else sv.nrm2_in[:, c],
                       None if obj is None else sv.att_in[:, c], obj, k, g_obj)
    t = sv.q_in[:, c] * sv.nrm_in[:, c].unsqueeze(-1)          # tanh output = q * nrm
    gu = gt * (1 - t * t)
    gx = gu @ W.W_leaf
    dW_leaf = torch.einsum('bnd,bne->de', gu, x)
    db_leaf = gu.sum((0, 1))

    # weight grads of the per-cell projections: one [PI*D, B*C] x [B*C, D] GEMM each  (kernel: gemm_tn)
    dWcat_in = torch.einsum('bcp,bcd->pd', GP_in, ih)
    dWcat_out = torch.einsum('bcp,bcd->pd', GP_out, oh)

    G = {}
    if W.share:
        dW1 = torch.cat([dWcat_in[0:D], dWcat_in[D:2 * D] + dWcat_out[0:D]], 1)
        G['inside_compose_func.h_fcs.0.weight'] = dW1
        G['inside_compose_func.h_fcs.0.bias'] = db1 + dob1
        G['inside_compose_func.h_fcs.2.weight'] = dW2 + doW2
        G['inside_compose_func.h_fcs.2.bias'] = db2 + dob2
        G['inside_score_func.mat'] = dWcat_in[2 * D:3 * D] + dWcat_out[D:2 * D]
    else:
        G['inside_compose_func.h_fcs.0.weight'] = torch.cat([dWcat_in[0:D], dWcat_in[D:2 * D]], 1)
        G['inside_compose_func.h_fcs.0.bias'] = db1
        G['inside_compose_func.h_fcs.2.weight'] = dW2
        G['inside_compose_func.h_fcs.2.bias'] = db2
        G['inside_score_func.mat'] = dWcat_in[2 * D:3 * D]
        G['outside_compose_func.h_fcs.0.weight'] = torch.cat([dWcat_in[3 * D:4 * D], dWcat_out[0:D]], 1)
        G['outside_compose_func.h_fcs.0.bias'] = dob1
        G['outside_compose_func.h_fcs.2.weight'] = doW2
        G['outside_compose_func.h_fcs.2.bias'] = dob2
        G['outside_score_func.mat'] = dWcat_out[D:2 * D]
    G['inside_compose_func.leaf_fc.weight'] = dW_leaf
    G['inside_compose_func.leaf_fc.bias'] = db_leaf
    G['root_vector_out_h'] = g_root
    return G, gx, g_obj
