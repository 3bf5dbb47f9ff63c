"""Golden vectors for the whole training step, produced by the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_step.py

``cliora.net.trainer.build_net`` builds Embed + ImageEncoder + CLIORA DioraMLP + the three losses + Adam exactly
as ``scripts/train.py`` does (CPU, --obj_feats --vg_loss --use_contr, --emb none); ``Trainer.step`` is then run
for three batches.  The only intervention is the one make_golden.py already uses: the AttentionHead's
``nn.Dropout`` is replaced by a module that replays a stored keep-mask, so the step is reproducible without the
reference's RNG stream.  Stored: initial state_dict, the batches, per-step losses and the state_dict after the
last step (i.e. after three clip(5.0)+Adam updates).  Writes tests/golden/train_step.pt.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, '/root/reference')
import cliora.net.trainer as ref_trainer  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import MaskQueueDropout  # noqa: E402


def main():
    torch.manual_seed(11)
    D, V, E, K, R, F = 32, 60, 24, 9, 5, 2048
    opts = argparse.Namespace(lr=2e-3, hidden_dim=D, k_neg=K, margin=1.0, vl_margin=0.2, hinge_margin=1.0,
                              alpha_contr=0.8, alpha_vg=0.6, normalize='unit', cuda=False, local_rank=0, share=True,
                              arch='mlp', obj_feats=True, multigpu=False, emb='none', vg_loss=True, use_contr=True,
                              use_contr_ce=False, visualize=False, load_model_path=None, experiment_name='golden')
    emb = torch.nn.Embedding(V, E)
    trainer = ref_trainer.build_net(opts, emb)
    net = trainer.net
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():       # the reference zero-initialises ImageEncoder: draw it so the visual path is live
        for p in net.image_encoder.parameters() if hasattr(net, 'image_encoder') else net.img_encoder.parameters():
            p.copy_(0.05 * torch.randn(p.shape, generator=g))
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    steps = []
    for (B, n) in [(4, 6), (3, 4), (4, 6)]:
        C = n * (n + 1) // 2
        keep = torch.rand(B, C, R, generator=g) >= 0.1
        net.diora.atten_head.dropout = MaskQueueDropout(keep, n)
        batch = dict(example_ids=list(range(B)), sentences=torch.randint(0, V, (B, n), generator=g),
                     image_feats=torch.zeros(B, 1), neg_samples=torch.randperm(V, generator=g)[:K],
                     obj_feats=torch.rand(B, R, F, generator=g), boxes=torch.zeros(B, R, 4),
                     obj_cates=torch.zeros(B, R, dtype=torch.int64), GT=[[(0, n - 1)]] * B, batch_size=B, length=n)
        res = trainer.step(batch, train=True)
        steps.append(dict(batch={k: v for k, v in batch.items() if torch.is_tensor(v)}, keep=keep,
                          result={k: v for k, v in res.items() if 'loss' in k}))
    final = {k: v.detach().clone() for k, v in net.state_dict().items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'train_step.pt')
    torch.save(dict(D=D, V=V, E=E, K=K, R=R, F=F, alpha_vg=0.6, alpha_contr=0.8, vl_margin=0.2, lr=2e-3, init=init,
                    steps=steps, final=final), path)
    print('wrote', path, os.path.getsize(path), 'bytes')
    for s in steps:
        print(s['result'])
    print(sorted(init.keys()))


if __name__ == '__main__':
    main()
