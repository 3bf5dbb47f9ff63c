"""Whole training step through the public API: eager vs CUDA-graph replay, and against the CPU oracle step."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(D=64, V=200, E=32, k_neg=10, seed=5):
    from cliora_b200.net.trainer import build_net
    torch.manual_seed(seed)
    opts = argparse.Namespace(arch='mlp', hidden_dim=D, k_neg=k_neg, margin=1.0, vl_margin=0.2, alpha_contr=1.0,
                              alpha_vg=1.0, vg_loss=True, use_contr=True, use_contr_ce=False, obj_feats=True,
                              normalize='unit', share=True, cuda=True, lr=2e-3)
    tr = build_net(opts, torch.nn.Embedding(V, E))
    with torch.no_grad():
        for p in tr.net.img_encoder.parameters():
            p.normal_(0, 0.02)
    return tr


def _reset_optimizer(tr):
    if hasattr(tr.optimizer, 'reset_state'):
        tr.optimizer.reset_state()
        return
    for st in tr.optimizer.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()


def _batch(B=6, n=7, V=200, R=9, F=2048, k_neg=10, seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(sentences=torch.randint(0, V, (B, n), generator=g).cuda(),
                neg_samples=torch.randperm(V, generator=g)[:k_neg].cuda(),
                obj_feats=torch.rand(B, R, F, generator=g).cuda(), batch_size=B, length=n)


@pytest.mark.parametrize('B,n,R', [(6, 7, 9), (2, 1, 3), (1, 2, 1), (5, 3, 64)])
def test_graphed_step_matches_eager_step(B, n, R):
    """Same weights, same batches, dropout off: CUDA-graph replay must reproduce the eager losses and weights
    (also for single-word / single-sentence / single-region batches)."""
    batches = [_batch(B=B, n=n, R=R, seed=i) for i in range(4)]
    eager, graphed = _trainer(), _trainer()
    for tr in (eager, graphed):
        tr.net.diora.atten_head.dropout.p = 0.0
    sd = {k: v.clone() for k, v in eager.net.state_dict().items()}
    graphed.capture(batches[0], warmup=2)          # warm-up steps update the weights and Adam state ...
    graphed.net.load_state_dict(sd)                # ... so reset both before comparing
    _reset_optimizer(graphed)
    la = [eager.step(x, train=True, sync_result=False)['total_loss'].item() for x in batches]
    lb = [graphed.step_graphed(x).item() for x in batches]
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-4 * abs(x), (la, lb)
    for (k, p), (_, q) in zip(eager.net.named_parameters(), graphed.net.named_parameters()):
        # Adam divides by sqrt(v): near-zero gradients turn fp noise into +-lr moves, so only gross errors are checked
        assert torch.allclose(p, q, rtol=1e-2, atol=1e-2), k


def _load_cpu_weights(net, cpu):
    with torch.no_grad():
        sd = net.diora.state_dict()
        for k in sd:
            sd[k].copy_(cpu.P[k if k in cpu.P else k.replace('outside_', 'inside_')])
        net.embed.embeddings.weight.copy_(cpu.emb)
        net.embed.mat.copy_(cpu.mat); net.embed.mat1.copy_(cpu.mat1)
        net.reconstruct_softmax_loss.mat.copy_(cpu.recon_mat)
        net.img_encoder.fc.weight.copy_(cpu.enc['fc.weight']); net.img_encoder.fc.bias.copy_(cpu.enc['fc.bias'])
        net.img_encoder.fc_vis.weight.copy_(cpu.enc['fc_vis.weight']); net.img_encoder.fc_vis.bias.copy_(cpu.enc['fc_vis.bias'])


def _oracle_step_grads(cpu, bt, keep):
    total, parts = cpu.loss(bt['sentences'].cpu(), bt['neg_samples'].cpu(), bt['obj_feats'].cpu(), keep)
    total.backward()
    g = {'embed.mat': cpu.mat.grad, 'embed.mat1': cpu.mat1.grad, 'reconstruct_softmax_loss.mat': cpu.recon_mat.grad}
    for k, v in cpu.enc.items():
        g['img_encoder.' + k] = v.grad
    for k, v in cpu.P.items():
        if not k.startswith('outside_'):
            g['diora.' + k] = v.grad
    return [p.item() for p in parts], g


def _check_step_vs_cpu_oracle(B, n, R, D, K, V=200, E=32, chains=None, gtol=2e-4, arbiter64=False):
    """``arbiter64``: gradients are compared with the oracle step in float64; where the float32 oracle (the
    reference's own precision) is itself further than ``gtol`` from float64 the bound is 3x that distance (at the
    bench workload the reference in float32 is 5e-4 of max away from float64 on the deepest gradients: 51 million
    ReLU units, a handful of which sit within rounding of their kink in any one batch)."""
    from conftest import rel_err
    from oracle.cliora_oracle import CpuClioraStep
    F = 2048
    tr = _trainer(D, V, E, K)
    cpu = CpuClioraStep(D=D, E=E, V=V, F=F, k_neg=K, seed=3)
    net = tr.net
    net.diora.chains = chains
    _load_cpu_weights(net, cpu)
    bt = _batch(B=B, n=n, V=V, R=R, k_neg=K)
    keep = torch.rand(B, n * (n + 1) // 2, bt['obj_feats'].shape[1]) >= 0.1
    net.train()
    net.diora.set_dropout_mask(keep.cuda())
    out = tr.run_net(bt, None, compute_loss=True)
    parts32, g32 = _oracle_step_grads(cpu, bt, keep)
    if arbiter64:
        cpu64 = CpuClioraStep(D=D, E=E, V=V, F=F, k_neg=K, seed=3, dtype=torch.float64)
        parts, ref = _oracle_step_grads(cpu64, bt, keep)
    else:
        parts, ref = parts32, g32
    mine = out['total_loss'].view(-1).cpu()
    for x, y in zip(mine.tolist(), parts):
        assert abs(x - y) <= 1e-4 * max(abs(y), 1e-3), (mine, parts)
    out['total_loss'].sum().backward()
    named = dict(net.named_parameters())
    checked = 0
    for k, g in ref.items():
        if k == 'img_encoder.fc_vis.bias':      # zero true gradient (the VG cross-entropy is shift-invariant): fp noise only
            continue
        if g is None:      # n == 1: the compose MLP is never used
            assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, k
            continue
        floor = rel_err(g32[k], g) if arbiter64 else 0.0
        assert rel_err(named[k].grad, g) < max(gtol, (3 if arbiter64 else 2) * floor), (k, floor)
        checked += 1
    assert checked >= 7
    return tr, cpu


@pytest.mark.parametrize('B,n,R,D,K', [(6, 7, 9, 64, 10), (1, 2, 1, 64, 1), (2, 1, 3, 36, 5), (9, 8, 36, 132, 100),
                                       (3, 3, 64, 32, 2)])
def test_train_step_losses_vs_cpu_oracle_step(B, n, R, D, K):
    """Net.forward + three losses through the fused product path == the oracle's dense CPU step (first step),
    including degenerate batches (one sentence, one word, one region, one negative)."""
    _check_step_vs_cpu_oracle(B, n, R, D, K)


def test_bench_workload_step_vs_cpu_oracle_step():
    """The configuration bench.py times (BASELINE.json config[1]): B=32, n=20, D=400, R=36, F=2048, V=8000,
    E=1024, 100 negatives, two sentence chains.  Losses 1e-4, every gradient 2e-4 of max against the dense CPU
    oracle step on the same batch and dropout mask; then the same step under CUDA-graph replay (dropout off,
    since a replayed graph draws its own mask) reproduces the oracle's loss."""
    from oracle.cliora_oracle import CpuClioraStep
    tr, cpu = _check_step_vs_cpu_oracle(32, 20, 36, 400, 100, V=8000, E=1024, chains=2, arbiter64=True)
    # graph replay of the very same configuration (chains = 2), dropout off on both sides
    tr2 = _trainer(400, 8000, 1024, 100)
    tr2.net.diora.chains = 2
    tr2.net.diora.atten_head.dropout.p = 0.0
    bt = _batch(B=32, n=20, V=8000, R=36, k_neg=100, seed=4)
    tr2.capture(bt, warmup=1)
    cpu2 = CpuClioraStep(D=400, E=1024, V=8000, F=2048, k_neg=100, seed=3)
    _load_cpu_weights(tr2.net, cpu2)
    if hasattr(tr2.optimizer, 'reset_state'):
        tr2.optimizer.reset_state()
    loss = tr2.step_graphed(bt).item()
    total, _ = cpu2.loss(bt['sentences'].cpu(), bt['neg_samples'].cpu(), bt['obj_feats'].cpu(), None)
    assert abs(loss - total.item()) <= 1e-4 * abs(total.item()), (loss, total.item())


def test_c5_batch128_four_chains_vs_cpu_oracle():
    """config[4] shape per GPU: batch 128, length 20, four sentence chains.  The oracle (dense, CPU) runs on a
    16-sentence slice; sentences are independent in the chart, so the slice's chart must match the same rows of
    the full-batch GPU run."""
    from conftest import rel_err
    from oracle import cliora_oracle as O
    from cliora_b200.net.cliora import DioraMLP
    from test_gpu_chart import _fill
    B, n, D, R = 128, 20, 400, 36
    P0 = O.init_params(D, share=True, seed=7)
    g = torch.Generator().manual_seed(15)
    x = torch.randn(B, n, D, generator=g)
    obj = 0.05 * torch.randn(B, R, D, generator=g)
    keep = torch.rand(B, O.num_cells(n), R, generator=g) >= 0.1
    m = DioraMLP(D).cuda()
    m.chains = 4
    _fill(m, P0)
    m.train()
    m.set_dropout_mask(keep.cuda())
    with torch.no_grad():
        m(x.cuda(), x.cuda(), obj.cuda(), obj.cuda())
    sl = slice(56, 72)       # straddles the boundary between chains 1 and 2
    out = O.chart_forward(P0, x[sl], obj[sl], keep[sl])
    for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s'):
        assert rel_err(getattr(m, k)[sl], getattr(out, k)) < 1e-4, k


def test_split_graph_path_used_for_data_parallel():
    """With a grad_sync hook the step is two graphs around an eager all-reduce; same losses as eager."""
    batches = [_batch(seed=i) for i in range(3)]
    eager, graphed = _trainer(), _trainer()
    calls = []
    graphed.grad_sync = lambda: calls.append(1)
    for tr in (eager, graphed):
        tr.net.diora.atten_head.dropout.p = 0.0
    sd = {k: v.clone() for k, v in eager.net.state_dict().items()}
    graphed.capture(batches[0], warmup=1)
    assert graphed._graph_opt is not None
    graphed.net.load_state_dict(sd)
    _reset_optimizer(graphed)
    n0 = len(calls)
    la = [eager.step(x, train=True, sync_result=False)['total_loss'].item() for x in batches]
    lb = [graphed.step_graphed(x).item() for x in batches]
    assert len(calls) - n0 == len(batches)
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-4 * abs(x), (la, lb)


def test_prefetched_batches_give_same_losses():
    """Trainer.prefetch (side-stream H2D into staging slots) feeds step_graphed the right batch every time."""
    batches = [{k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in _batch(seed=i).items()} for i in range(5)]
    a, b = _trainer(), _trainer()
    for tr in (a, b):
        tr.net.diora.atten_head.dropout.p = 0.0
    sd = {k: v.clone() for k, v in a.net.state_dict().items()}
    for tr in (a, b):
        tr.capture(batches[0], warmup=1)
        tr.net.load_state_dict(sd)
        _reset_optimizer(tr)
    la = [a.step_graphed(x).item() for x in batches]
    h = b.prefetch(batches[0])
    lb = []
    for i in range(len(batches)):
        nxt = b.prefetch(batches[(i + 1) % len(batches)])
        lb.append(b.step_graphed(h).item())
        h = nxt
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-4 * abs(x), (la, lb)


def test_fused_clip_adam_matches_torch():
    """FusedClipAdam == clip_grad_norm_(5.0) + torch Adam (trainer.py:450-455,580) over several steps."""
    from cliora_b200.optim import FusedClipAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(400, 800), (400,), (400, 400), (7,), (1025, 3)]
    pa = [torch.randn(*s, generator=g).cuda().requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    fa = FusedClipAdam(pa, lr=2e-3)
    fb = torch.optim.Adam(pb, lr=2e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(4):
        scale = 10.0 if step % 2 == 0 else 0.01          # exercise both the clipping and the no-clip branch
        grads = [scale * torch.randn(*s, generator=g).cuda() for s in shapes]
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(pb, 5.0)
        fb.step()
        fa.step()
        assert abs(fa.grad_norm.item() - norm.item()) <= 1e-5 * norm.item()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)


def test_torch_adam_path_still_works():
    """fused=False keeps torch's capturable Adam + clip_grad_norm_ (also inside the step graph)."""
    import torch.optim as optim
    tr = _trainer()
    tr.init_optimizer(optim.Adam, dict(lr=2e-3, betas=(0.9, 0.999), eps=1e-8), fused=False)
    tr.net.diora.atten_head.dropout.p = 0.0
    b = _batch(seed=1)
    tr.capture(b, warmup=1)
    l1 = tr.step_graphed(b).item()
    l2 = tr.step_graphed(b).item()
    assert l2 < l1          # same batch twice: the update must reduce its loss


@pytest.mark.parametrize('split', [False, True])
def test_step_auto_replays_per_shape_graphs(split):
    """A stream of batches with varying (batch, length): step_auto runs a shape eagerly, captures it on its
    second occurrence and replays afterwards; losses follow the eager trainer step for step.  ``split`` puts a
    (dummy) gradient-sync hook in, i.e. the data-parallel two-graph form with per-graph gradient tensors."""
    shapes = [(6, 7), (4, 5), (6, 3)]
    order = [0, 1, 0, 2, 0, 1, 1, 2, 0, 2, 1, 0]
    batches = [_batch(B=shapes[s][0], n=shapes[s][1], seed=i) for i, s in enumerate(order)]
    eager, auto = _trainer(), _trainer()
    calls = []
    if split:
        auto.grad_sync = lambda: calls.append(1)
    for tr in (eager, auto):
        tr.net.diora.atten_head.dropout.p = 0.0
    auto.net.load_state_dict({k: v.clone() for k, v in eager.net.state_dict().items()})
    la = [eager.step(x, train=True, sync_result=False)['total_loss'].item() for x in batches]
    lb = [auto.step_auto(x).item() for x in batches]
    assert len(auto._graphs) == 3
    for i, (x, y) in enumerate(zip(la, lb)):
        # Adam amplifies fp noise on near-zero gradients, so the trajectories drift apart slowly
        assert abs(x - y) <= (2e-4 + 2e-3 * i) * abs(x), (i, la, lb)
    if split:
        assert len(calls) == len(batches)      # exactly one gradient sync per step (none during a capture)


def test_product_trainer_matches_reference_trainer_golden():
    """tests/golden/train_step.pt: three Trainer.step calls of the UNMODIFIED reference (build_net, CPU).  The
    product's build_net loads the reference's state_dict as is (same keys) and must reproduce the losses of all
    three steps (steps 2 and 3 depend on the clip+Adam updates) and the updated weights."""
    import os
    from cliora_b200.net.trainer import build_net
    g = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'train_step.pt'), weights_only=False)
    opts = argparse.Namespace(arch='mlp', hidden_dim=g['D'], k_neg=g['K'], margin=1.0, vl_margin=g['vl_margin'],
                              alpha_contr=g['alpha_contr'], alpha_vg=g['alpha_vg'], vg_loss=True, use_contr=True,
                              use_contr_ce=False, obj_feats=True, normalize='unit', share=True, cuda=True,
                              lr=g['lr'])
    tr = build_net(opts, torch.nn.Embedding(g['V'], g['E']))
    tr.net.load_state_dict(g['init'], strict=True)
    for i, s in enumerate(g['steps']):
        batch = {k: v.cuda() for k, v in s['batch'].items()}
        batch.update(batch_size=batch['sentences'].shape[0], length=batch['sentences'].shape[1])
        tr.net.train()
        tr.net.diora.set_dropout_mask(s['keep'].cuda())
        res = tr.step(batch, train=True)
        for k, v in s['result'].items():
            assert res[k] == pytest.approx(v, rel=5e-4, abs=1e-5), (i, k, res, s['result'])
    lr = g['lr']
    sd = tr.net.state_dict()
    for k, ref in g['final'].items():
        if k == 'img_encoder.fc_vis.bias':      # zero true gradient: Adam amplifies rounding noise (see the oracle test)
            continue
        d = (sd[k].cpu() - ref).abs()
        assert float(d.max()) <= 3 * 2 * lr + 1e-6, k     # hard bound: at most lr per step per element, both ways
        assert float(d.mean()) <= 0.1 * lr, k
