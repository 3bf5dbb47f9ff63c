"""Pins the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU-only."""
import os

import pytest
import torch

from oracle import cliora_oracle as O
from conftest import rel_err

TOL = 2e-5   # fp32 vs fp32, different op order (reference self-noise is ~5e-7 of max)


def test_index_closed_forms(golden):
    blob = golden('index.pt')
    for n in range(2, 13):
        assert blob[('offset', n)] == {l: o for l, o in enumerate(O.level_offsets(n))}
        for level in range(1, n):
            l, r = O.inside_pairs(n, level)
            gl, gr = blob[('inside', n, level)]
            assert torch.equal(l, gl) and torch.equal(r, gr)
        for level in range(0, n - 1):
            p, s = O.outside_pairs(n, level)
            gp, gs = blob[('outside', n, level)]
            assert torch.equal(p, gp) and torch.equal(s, gs)


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference not mounted')
def test_index_closed_forms_live_up_to_40():
    import sys
    sys.path.insert(0, '/root/reference')
    from cliora.net.inside_index import get_inside_index
    from cliora.net.outside_index import get_outside_index
    for n in (13, 20, 31, 40):
        for level in range(1, n):
            l, r = O.inside_pairs(n, level)
            gl, gr = get_inside_index(n, level)
            assert torch.equal(l, gl) and torch.equal(r, gr)
        for level in range(0, n - 1):
            p, s = O.outside_pairs(n, level)
            gp, gs = get_outside_index(n, level)
            assert torch.equal(p, gp) and torch.equal(s, gs)


def _params(blob):
    if 'params' in blob:
        return {k: v.clone().requires_grad_() for k, v in blob['params'].items()}
    P = O.init_params(blob['D'], share=blob['share'], seed=blob['seed'])
    return {k: v.clone().requires_grad_() for k, v in P.items()}


def _share_alias(P, share):
    """The reference aliases outside_* to inside_* modules when share=True."""
    if share:
        for k in list(P):
            if k.startswith('inside_'):
                P['outside_' + k[len('inside_'):]] = P[k]
    return P


DIORA = ['diora_b2_n5_d16_share.pt', 'diora_b3_n7_d32_noshare.pt', 'diora_b2_n2_d16_share.pt',
         'diora_b1_n1_d16_share.pt', 'diora_b2_n6_d400_share.pt']


@pytest.mark.parametrize('name', DIORA)
def test_diora_forward_backward(golden, name):
    blob = golden(name)
    P = _share_alias(_params(blob), blob['share'])
    x = blob['x'].clone().requires_grad_()
    out = O.chart_forward(P, x)
    for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s'):
        assert rel_err(getattr(out, k), blob[k]) < TOL, k
    loss = sum((getattr(out, k) * blob['g_' + k]).sum() for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s'))
    loss.backward()
    assert rel_err(x.grad, blob['grad_x']) < 1e-4
    st = blob.get('grads_strided')
    for k, g in blob['grads'].items():
        if blob['share'] and k.startswith('outside_'):
            continue
        mine = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        if st and mine.dim() == 2:
            mine = mine[::st, ::st]
        if g.abs().max() == 0:
            assert mine.abs().max() == 0, k
        else:
            assert rel_err(mine, g) < 1e-4, k


CLIORA = ['cliora_b3_n6_d32_r5_eval.pt', 'cliora_b3_n6_d32_r5_train.pt', 'cliora_b4_n9_d48_r36_train.pt']


@pytest.mark.parametrize('name', CLIORA)
def test_cliora_forward_losses_backward(golden, name):
    blob = golden(name)
    P = _share_alias({k: v.clone().requires_grad_() for k, v in blob['params'].items()}, True)
    leafs = {k: blob[k].clone().requires_grad_() for k in ('x_span', 'x_word', 'obj_span', 'obj_word')}
    keep = blob['keep'] if blob['train'] else None
    out = O.chart_forward(P, leafs['x_span'], leafs['obj_span'], keep)
    for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s'):
        assert rel_err(getattr(out, k), blob[k]) < TOL, k
    aas = O.all_atten_score(out.inside_h, out.outside_h, leafs['obj_span'])
    vg = O.vg_atten_score(leafs['x_word'], leafs['obj_word'], training=blob['train'], all_atten=aas)
    assert rel_err(aas, blob['all_atten_score']) < TOL
    assert rel_err(vg, blob['vg_atten_score']) < TOL
    assert rel_err(O.atten_score(vg), blob['atten_score']) < TOL
    mat = blob['recon_mat'].clone().requires_grad_()
    l_rec = O.reconstruction_loss(blob['emb_weight'], mat, blob['sentences'], blob['neg_samples'], out.outside_h)
    l_vg = O.vg_loss(vg, blob['alpha_vg'])
    l_con = O.contrastive_loss(aas, out.inside_s, out.outside_s, blob['margin'], blob['alpha_contr'])
    assert abs(l_rec.item() - blob['loss_recon'].item()) <= 1e-5 * abs(blob['loss_recon'].item())
    assert abs(l_vg.item() - blob['loss_vg'].item()) <= 1e-5 * abs(blob['loss_vg'].item())
    assert abs(l_con.item() - blob['loss_contr'].item()) <= 1e-5 * abs(blob['loss_contr'].item())
    (l_rec + l_vg + l_con).backward()
    assert rel_err(mat.grad, blob['grad_recon_mat']) < 1e-4
    for k in ('x_span', 'x_word', 'obj_span', 'obj_word'):
        assert rel_err(leafs[k].grad, blob['grad_' + k]) < 1e-4, k
    for k, g in blob['grads'].items():
        if k.startswith('outside_'):
            continue
        mine = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        assert rel_err(mine, g) < 1e-4, k


def test_cky_trees(golden):
    blob = golden('cky_b6_n9_d24.pt')
    trees = O.cky_trees(blob['split_scores'], blob['B'], blob['n'])
    assert trees == blob['trees']
    # and the oracle's own forward reproduces the split scores the reference hook saw
    P = _share_alias({k: v.clone() for k, v in blob['params'].items()}, True)
    out = O.chart_forward(P, blob['x'])
    for level, s in blob['split_scores'].items():
        assert rel_err(out.split_scores[level], s) < TOL
    assert O.cky_trees(out.split_scores, blob['B'], blob['n']) == blob['trees']


@pytest.mark.skipif(not os.path.isdir('/root/reference/cliora'), reason='needs the reference checkout')
@pytest.mark.parametrize('B,n,D,R,share,seed', [(2, 3, 8, 0, True, 1), (1, 12, 20, 0, False, 2), (3, 8, 36, 7, True, 3),
                                                (2, 10, 16, 36, True, 4), (5, 4, 12, 1, True, 5),
                                                (2, 16, 24, 0, True, 6)])
def test_oracle_against_the_imported_reference_live(B, n, D, R, share, seed):
    """Beyond the committed fixtures: the oracle next to the unmodified reference modules (imported from
    /root/reference, eval mode) on further shapes - chart tensors, alignment tensors and every gradient."""
    import sys
    sys.path.insert(0, '/root/reference')
    if R:
        from cliora.net.cliora import DioraMLP
        m = DioraMLP(D, outside=True, normalize='unit', compress=False, share=share)
    else:
        from cliora.net.diora import DioraMLP
        m = DioraMLP(D, outside=True, normalize='unit', compress=False, share=share)
    torch.manual_seed(seed)
    m.reset_parameters() if hasattr(m, 'reset_parameters') else None
    m.eval()
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.randn(B, n, D, generator=g, requires_grad=True)
    xw = torch.randn(B, n, D, generator=g, requires_grad=True)
    obj = (0.3 * torch.randn(B, max(R, 1), D, generator=g)).requires_grad_()
    objw = (0.3 * torch.randn(B, max(R, 1), D, generator=g)).requires_grad_()
    if R:
        m(x, xw, obj, objw)
    else:
        m(x, xw)
    names = ('inside_h', 'inside_s', 'outside_h', 'outside_s')
    ct = {k: torch.randn(getattr(m, k).shape, generator=g) for k in names}
    loss = sum((getattr(m, k) * ct[k]).sum() for k in names)
    if R:
        loss = loss + (m.all_atten_score * 0.01).sum() + (m.vg_atten_score * 0.01).sum()
    loss.backward()

    P = _share_alias({k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()
                      if not (share and k.startswith('outside_'))}, share)
    x2, xw2 = x.detach().clone().requires_grad_(), xw.detach().clone().requires_grad_()
    obj2, objw2 = obj.detach().clone().requires_grad_(), objw.detach().clone().requires_grad_()
    out = O.chart_forward(P, x2, obj2 if R else None, None)
    for k in names:
        assert rel_err(getattr(out, k), getattr(m, k)) < TOL, k
    loss2 = sum((getattr(out, k) * ct[k]).sum() for k in names)
    if R:
        aas = O.all_atten_score(out.inside_h, out.outside_h, obj2)
        vg = O.vg_atten_score(xw2, objw2, training=False, all_atten=aas)
        assert rel_err(aas, m.all_atten_score) < TOL and rel_err(vg, m.vg_atten_score) < TOL
        loss2 = loss2 + (aas * 0.01).sum() + (vg * 0.01).sum()
    loss2.backward()
    assert rel_err(x2.grad, x.grad) < 1e-4
    if R:
        assert rel_err(obj2.grad, obj.grad) < 1e-4 and rel_err(objw2.grad, objw.grad) < 1e-4
        assert rel_err(xw2.grad, xw.grad) < 1e-4
    for k, p in m.named_parameters():
        if share and k.startswith('outside_'):
            continue
        if p.grad is None:
            assert P[k].grad is None or float(P[k].grad.abs().max()) == 0.0, k
        else:
            assert rel_err(P[k].grad, p.grad) < 1e-4, k
