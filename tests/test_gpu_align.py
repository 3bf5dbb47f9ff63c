"""Alignment kernels, fused losses and the whole CLIORA loss/gradient against golden fixtures + oracle."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize('B,ncell,D,R', [(3, 7, 32, 5), (8, 105, 400, 36), (5, 20, 48, 64), (32, 105, 400, 36),
                                         (9, 50, 400, 17)])   # the last two take the tcgen05 path
def test_atten_max_fwd_bwd(B, ncell, D, R):
    from cliora_b200.net.cliora import AttenMax, AttenScores
    g = torch.Generator().manual_seed(B * ncell)
    h = torch.randn(B, ncell, D, generator=g).cuda().requires_grad_()
    obj = torch.randn(B, R, D, generator=g).cuda().requires_grad_()
    full = torch.einsum('abx,cdx->acbd', h.double(), obj.double())
    assert rel_err(AttenScores.apply(h, obj), full) < 2e-6
    ref_max, ref_arg = full.max(-1)
    smax, amax = AttenMax.apply(h, obj)
    assert rel_err(smax, ref_max) < 2e-6
    # argmax may differ only where the top two scores are within fp32 noise
    top2 = full.topk(2, -1).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-4
    assert torch.equal(amax.long()[clear], ref_arg[clear])
    w = torch.randn(B, B, ncell, generator=g).cuda()
    w[w.abs() < 0.3] = 0
    (smax * w).sum().backward()
    hd, od = h.detach().double().requires_grad_(), obj.detach().double().requires_grad_()
    (torch.einsum('abx,cdx->acbd', hd, od).max(-1).values * w.double()).sum().backward()
    assert rel_err(h.grad, hd.grad) < 2e-6
    assert rel_err(obj.grad, od.grad) < 2e-6


@pytest.mark.parametrize('B,n,R', [(3, 6, 5), (32, 20, 36), (5, 1, 4)])
def test_contrastive_and_vg_loss_kernels(B, n, R):
    from oracle import cliora_oracle as O
    from cliora_b200.net.losses import ContrastiveFn, VGLossFn
    g = torch.Generator().manual_seed(n)
    C = n * (n + 1) // 2
    aas = torch.randn(B, B, C, R, generator=g)
    ins = (0.3 * torch.randn(B, C, 1, generator=g)).requires_grad_()
    outs = (0.3 * torch.randn(B, C, 1, generator=g)).requires_grad_()
    aas.requires_grad_()
    ref = O.contrastive_loss(aas, ins, outs, margin=0.2, alpha=0.9)
    ref.backward()
    smax = aas.detach().max(-1).values[:, :, :C // 2].contiguous().cuda().requires_grad_()
    ic, oc = ins.detach().squeeze(-1).cuda().requires_grad_(), outs.detach().squeeze(-1).cuda().requires_grad_()
    loss = ContrastiveFn.apply(smax, ic, oc, 0.2, 0.9)
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(abs(ref.item()), 1e-6)
    loss.backward()
    if C // 2 > 0:
        # grad wrt smax == grad of the oracle wrt the max entries of all_atten_score
        ref_gs = aas.grad.sum(-1)[:, :, :C // 2]
        assert rel_err(smax.grad, ref_gs) < 1e-5
        assert rel_err(ic.grad, ins.grad.squeeze(-1)) < 1e-5
        assert rel_err(oc.grad, outs.grad.squeeze(-1)) < 1e-5
    vg = torch.randn(B, B, n, R, generator=g).requires_grad_()
    refv = O.vg_loss(vg, alpha=0.7)
    refv.backward()
    wmax = vg.detach().max(-1).values.contiguous().cuda().requires_grad_()
    lv = VGLossFn.apply(wmax, 0.7)
    assert abs(lv.item() - refv.item()) <= 1e-5 * abs(refv.item())
    lv.backward()
    assert rel_err(wmax.grad, vg.grad.sum(-1)) < 1e-5


CLIORA = ['cliora_b3_n6_d32_r5_eval.pt', 'cliora_b3_n6_d32_r5_train.pt', 'cliora_b4_n9_d48_r36_train.pt']


@pytest.mark.parametrize('name', CLIORA)
def test_cliora_losses_and_grads_vs_golden(golden, name):
    """Fused product path (no [B,B,cells,R] tensor): three losses and every gradient vs the reference."""
    from cliora_b200.net.cliora import DioraMLP
    from cliora_b200.net import trainer as T
    from test_gpu_chart import _fill, _grads
    blob = golden(name)
    m = DioraMLP(blob['D']).cuda()
    _fill(m, blob['params'])
    m.train() if blob['train'] else m.eval()
    if blob['train']:
        m.set_dropout_mask(blob['keep'].cuda())
    leaf = {k: blob[k].cuda().requires_grad_() for k in ('x_span', 'x_word', 'obj_span', 'obj_word')}
    m(leaf['x_span'], leaf['x_word'], leaf['obj_span'], leaf['obj_word'])
    emb = torch.nn.Embedding(blob['V'], blob['E']).cuda()
    emb.weight.data.copy_(blob['emb_weight'])
    emb.weight.requires_grad = False
    recon = T.ReconstructionSoftmaxLoss(emb, input_size=blob['E'], size=blob['D'], k_neg=blob['K']).cuda()
    recon.mat.data.copy_(blob['recon_mat'])
    sent, neg = blob['sentences'].cuda(), blob['neg_samples'].cuda()
    l_rec, _ = recon(sent, neg, m, {})
    vg = m.word_region_max()[0] if blob['train'] else m.vg_atten_score
    l_vg, _ = T.VGLoss(blob['alpha_vg'])(sent, vg)
    l_con, _ = T.ContrastiveLoss(blob['margin'], blob['alpha_contr'])(sent, m)
    for mine, key in ((l_rec, 'loss_recon'), (l_vg, 'loss_vg'), (l_con, 'loss_contr')):
        assert abs(mine.item() - blob[key].item()) <= 1e-4 * abs(blob[key].item()), key
    (l_rec + l_vg + l_con).backward()
    assert rel_err(recon.mat.grad, blob['grad_recon_mat']) < TOL
    for k in ('x_span', 'x_word', 'obj_span', 'obj_word'):
        assert rel_err(leaf[k].grad, blob['grad_' + k]) < TOL, k
    mine = _grads(m)
    for k, gref in blob['grads'].items():
        assert rel_err(mine[k], gref) < TOL, k


@pytest.mark.parametrize('B,n,D,K', [(3, 5, 32, 7), (32, 20, 400, 100), (2, 1, 48, 33)])
def test_fused_reconstruction_ce(B, n, D, K):
    """recon_ce kernels == the oracle's reconstruction_loss (trainer.py:46-78), loss and all three gradients."""
    from oracle import cliora_oracle as O
    from cliora_b200.net.losses import ReconCEFn
    g = torch.Generator().manual_seed(B * n + K)
    V, E = 50 + K, 24
    emb = torch.randn(V, E, generator=g)
    mat = (0.2 * torch.randn(D, E, generator=g)).requires_grad_()
    sent = torch.randint(0, V, (B, n), generator=g)
    neg = torch.randperm(V, generator=g)[:K]
    C = n * (n + 1) // 2
    oh = torch.randn(B, C, D, generator=g).requires_grad_()
    ref = O.reconstruction_loss(emb, mat, sent, neg, oh)
    ref.backward()
    cell = oh.detach()[:, :n].cuda().requires_grad_()
    pos = (emb[sent] @ mat.detach().t()).cuda().requires_grad_()
    ngv = (emb[neg] @ mat.detach().t()).cuda().requires_grad_()
    loss = ReconCEFn.apply(cell.reshape(B * n, D), pos.reshape(B * n, D), ngv)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    (loss * 1.7).backward()
    assert rel_err(cell.grad, 1.7 * oh.grad[:, :n]) < 1e-5
    # chain the pos / neg grads back to `mat` to compare with the oracle's gradient of the projection matrix
    gmat = pos.grad.reshape(B * n, D).cpu().t() @ emb[sent].reshape(B * n, E) + ngv.grad.cpu().t() @ emb[neg]
    assert rel_err(gmat, 1.7 * mat.grad) < 1e-5
