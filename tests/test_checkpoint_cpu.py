"""Trainer.save_model / load_model (reference: cliora/net/trainer.py:382-435): same file format, interchangeable
with the reference in both directions.  Host logic only, so it runs without a GPU."""
import argparse
import os
import sys

import pytest
import torch

from cliora_b200.net.trainer import Trainer, build_net


def _opts(**kw):
    base = dict(lr=2e-3, hidden_dim=16, k_neg=5, margin=1.0, vl_margin=0.2, hinge_margin=1.0, alpha_contr=1.0,
                alpha_vg=1.0, normalize='unit', cuda=False, local_rank=0, share=True, arch='mlp', obj_feats=True,
                multigpu=False, emb='none', vg_loss=True, use_contr=True, use_contr_ce=False, visualize=False,
                load_model_path=None, experiment_name='t')
    base.update(kw)
    return argparse.Namespace(**base)


def _mine(seed, **kw):
    torch.manual_seed(seed)
    return build_net(_opts(**kw), torch.nn.Embedding(30, 12))


def test_roundtrip_without_embeddings(tmp_path):
    a, b = _mine(1), _mine(2)
    path = str(tmp_path / 'model.pt')
    a.save_model(False, path)
    saved = torch.load(path)['state_dict']
    assert saved and not any('embeddings' in k for k in saved)
    before = b.net.embed.embeddings.weight.clone()
    Trainer.load_model(False, b.net, path)
    for (k, x), (_, y) in zip(a.net.state_dict().items(), b.net.state_dict().items()):
        if 'embeddings' in k:
            continue
        assert torch.equal(x, y), k
    assert torch.equal(b.net.embed.embeddings.weight, before)        # kept from the target net


def test_module_prefix_unknown_keys_and_build_net_loading(tmp_path):
    a = _mine(3)
    sd = {'module.' + k: v for k, v in a.net.state_dict().items()}    # as written by a DDP-wrapped reference net
    sd['module.something.else'] = torch.zeros(1)
    path = str(tmp_path / 'ddp.pt')
    torch.save({'state_dict': sd}, path)
    b = _mine(4, load_model_path=path)                                # build_net(load_model_path=...) like train.py
    for (k, x), (_, y) in zip(a.net.state_dict().items(), b.net.state_dict().items()):
        assert torch.equal(x, y), k


def test_freeze_helpers_and_parameter_norm():
    t = _mine(5)
    n_all = t.parameter_norm(requires_grad=False)
    t.freeze_diora()
    assert not any(p.requires_grad for p in t.net.diora.parameters())
    assert 0 < t.parameter_norm() < n_all
    t.freeze_except_vis()
    assert {k for k, p in t.net.named_parameters() if p.requires_grad} == {'img_encoder.fc_vis.weight',
                                                                           'img_encoder.fc_vis.bias'}


@pytest.mark.skipif(not os.path.isdir('/root/reference/cliora'), reason='needs the reference checkout')
def test_checkpoints_interchange_with_the_reference(tmp_path):
    sys.path.insert(0, '/root/reference')
    import cliora.net.trainer as ref
    torch.manual_seed(6)
    theirs = ref.build_net(_opts(), torch.nn.Embedding(30, 12))
    mine = _mine(7)
    p1, p2 = str(tmp_path / 'ref.pt'), str(tmp_path / 'mine.pt')
    theirs.save_model(True, p1)
    Trainer.load_model(True, mine.net, p1)                            # reference file -> this implementation
    assert list(mine.net.state_dict().keys()) == list(theirs.net.state_dict().keys())
    for (k, x), (_, y) in zip(theirs.net.state_dict().items(), mine.net.state_dict().items()):
        assert torch.equal(x, y), k
    with torch.no_grad():
        for p in mine.net.parameters():
            p.add_(1.0)
    mine.save_model(True, p2)
    ref.Trainer.load_model(True, theirs.net, p2)                      # this implementation's file -> reference
    for (k, x), (_, y) in zip(mine.net.state_dict().items(), theirs.net.state_dict().items()):
        assert torch.equal(x, y), k


@pytest.mark.skipif(not os.path.isdir('/root/reference/cliora'), reason='needs the reference checkout')
def test_public_surface_matches_the_reference():
    """Every public name of the reference's trainer/net classes exists here, except the ones INTEGRATION.md lists
    as fused away or out of scope."""
    sys.path.insert(0, '/root/reference')
    import cliora.net.cliora as rc
    import cliora.net.diora as rd
    import cliora.net.trainer as rt
    import cliora_b200.net.cliora as mc
    import cliora_b200.net.diora as md
    import cliora_b200.net.trainer as mt
    import cliora_b200.net.utils as mu

    def pub(o):
        return {n for n in dir(o) if not n.startswith('_')}
    fused = {'initialize_outside_root', 'inside_func', 'inside_pass', 'leaf_transform', 'outside_func', 'outside_pass'}
    assert pub(rd.DioraMLP(16)) - pub(md.DioraMLP(16)) == fused
    assert pub(rc.DioraMLP(16)) - pub(mc.DioraMLP(16)) == fused
    for cls, allowed in [('Trainer', set()), ('Net', {'visualization'}), ('Embed', set()),
                         ('ReconstructionSoftmaxLoss', set()), ('ContrastiveLoss', set()), ('VGLoss', set())]:
        assert pub(getattr(rt, cls)) - pub(getattr(mt, cls)) == allowed, cls
    assert mu.ImageEncoder is mt.ImageEncoder
