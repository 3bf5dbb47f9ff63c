"""Driver for the UNMODIFIED reference staged under oracle/_ref (see make_ref.py).  BASELINE INFRASTRUCTURE ONLY:
imported by bench.py's reference / cpu_baseline / gpu_baseline legs and by tests; never by cliora_b200/.

It calls the reference through its own public API exactly as cliora/scripts/train.py does:
``cliora.net.trainer.build_net(options, embeddings)`` (trainer.py:504-582) and ``Trainer.step(batch_map)``
(trainer.py:483-501), with the options of train_cliora.sh (--arch mlp --obj_feats --use_contr --vg_loss --emb none).
The only intervention is the one SURVEY.md section 8(d) prescribes for every arm: ImageEncoder is re-drawn
N(0, 0.02) because the reference zero-initialises it (cliora/net/utils.py:45-50), which would make the visual path
vanish.
"""
import argparse
import os
import sys
import time

import torch

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'cliora', 'net'))


def _import():
    if not available():
        raise ImportError('oracle/_ref is not staged (run python oracle/make_ref.py where /root/reference exists)')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import cliora.net.trainer as ref_trainer
    return ref_trainer


def build_reference_trainer(cfg, cuda=False, seed=1234, lr=2e-3, obj_feats=True):
    ref_trainer = _import()
    torch.manual_seed(seed)
    opts = argparse.Namespace(lr=lr, hidden_dim=cfg['D'], k_neg=cfg['k_neg'], margin=1.0, vl_margin=0.2,
                              hinge_margin=1.0, alpha_contr=1.0, alpha_vg=1.0, normalize='unit', cuda=cuda,
                              local_rank=0, share=True, arch='mlp', obj_feats=obj_feats, multigpu=False, emb='none',
                              vg_loss=obj_feats, use_contr=obj_feats, use_contr_ce=False, visualize=False,
                              load_model_path=None, experiment_name='bench', master_addr='127.0.0.1',
                              master_port='29500')
    emb = torch.nn.Embedding(cfg['V'], cfg['E'])
    trainer = ref_trainer.build_net(opts, emb)
    net = trainer.net
    enc = net.image_encoder if hasattr(net, 'image_encoder') else getattr(net, 'img_encoder', None)
    if enc is not None:
        with torch.no_grad():
            for p in enc.parameters():
                p.normal_(0, 0.02)
    return trainer


def reference_batch(cfg, batch, device):
    """The dict scripts/train.py hands to Trainer.step (keys consumed at trainer.py:437-448)."""
    B, n, R = cfg['B'], cfg['n'], cfg['R']
    dev = torch.device(device)
    return dict(example_ids=list(range(B)), sentences=batch['sentences'].to(dev),
                image_feats=torch.zeros(B, 1, device=dev), neg_samples=batch['neg_samples'].to(dev),
                obj_feats=batch['obj_feats'].to(dev), boxes=torch.zeros(B, R, 4, device=dev),
                obj_cates=torch.zeros(B, R, dtype=torch.int64, device=dev), GT=[[(0, n - 1)]] * B,
                batch_size=B, length=n)


def time_reference(cfg, batches, steps, warmup, cuda=False, threads=None):
    """(sentences/s, ms/step, threads) of Trainer.step on the given batches; CUDA events when cuda else wall clock."""
    if threads:
        torch.set_num_threads(threads)
    trainer = build_reference_trainer(cfg, cuda=cuda)
    dev = 'cuda' if cuda else 'cpu'
    bms = [reference_batch(cfg, b, dev) for b in batches]
    for i in range(warmup):
        trainer.step(bms[i % len(bms)], train=True)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            trainer.step(bms[i % len(bms)], train=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / max(steps, 1)
    else:
        t0 = time.perf_counter()
        for i in range(steps):
            trainer.step(bms[i % len(bms)], train=True)
        ms = (time.perf_counter() - t0) * 1e3 / max(steps, 1)
    return cfg['B'] * 1e3 / ms, ms, torch.get_num_threads()
