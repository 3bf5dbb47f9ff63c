"""The oracle's whole-step restatement (CpuClioraStep: Embed, ImageEncoder, chart, three losses, clip(5.0), Adam)
against tests/golden/train_step.pt, which the UNMODIFIED reference produced through its own build_net +
Trainer.step (tests/golden/make_golden_step.py).  This is what pins the checker used by
tests/test_gpu_trainer.py and the CPU arm of bench.py."""
import os

import pytest
import torch

from oracle.cliora_oracle import CpuClioraStep

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'train_step.pt')


def _load_into(step, sd):
    with torch.no_grad():
        for k, v in step.P.items():
            v.copy_(sd['diora.' + k])
        step.emb.copy_(sd['embed.embeddings.weight'])
        step.mat.copy_(sd['embed.mat'])
        step.mat1.copy_(sd['embed.mat1'])
        step.recon_mat.copy_(sd['reconstruct_softmax_loss.mat'])
        for k in step.enc:
            step.enc[k].copy_(sd['img_encoder.' + k])


def test_oracle_step_matches_reference_trainer():
    g = torch.load(GOLD, weights_only=False)
    step = CpuClioraStep(D=g['D'], E=g['E'], V=g['V'], F=g['F'], k_neg=g['K'], lr=g['lr'], alpha_vg=g['alpha_vg'],
                         alpha_contr=g['alpha_contr'], margin=g['vl_margin'])
    _load_into(step, g['init'])
    for i, s in enumerate(g['steps']):
        b = s['batch']
        total, parts = step.step(b['sentences'], b['neg_samples'], b['obj_feats'], s['keep'])
        want = s['result']
        got = dict(reconstruction_softmax_loss=parts[0], vg_loss=parts[1], contrastive_loss=parts[2], total_loss=total)
        for k, v in want.items():
            # steps 2 and 3 run on weights the oracle's own clip+Adam produced, so they check the update too
            assert got[k] == pytest.approx(v, rel=2e-4, abs=1e-5), (i, k, got, want)
    fin = g['final']
    pairs = [('diora.' + k, step.P[k]) for k in step.P] + [
        ('embed.mat', step.mat), ('embed.mat1', step.mat1), ('reconstruct_softmax_loss.mat', step.recon_mat)] + [
        ('img_encoder.' + k, step.enc[k]) for k in step.enc]
    for name, a in pairs:
        if name == 'img_encoder.fc_vis.bias':
            # its true gradient is zero: the bias adds x_word . b to every image's logit of the grounding
            # cross-entropy, which cancels in the softmax over images.  What is left is rounding noise, and Adam
            # turns noise into +-lr steps, so two correct implementations disagree here by construction.
            continue
        moved = float((fin[name] - g['init'][name]).abs().mean())
        assert moved > 1e-4, name                                   # three real updates happened
        assert float((a.detach() - fin[name]).abs().max()) <= 5e-5, name
    assert torch.equal(step.emb, fin['embed.embeddings.weight'])      # frozen with --obj_feats (trainer.py:538-541)
