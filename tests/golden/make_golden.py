"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these
fixtures -- outputs of the reference itself on seeded inputs -- are what pins
the oracle (oracle/cliora_oracle.py) and, through it and directly, the CUDA
path.  Nothing here is imported by the product.
"""
import os
import sys
import types

import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from cliora.net import diora as ref_diora            # noqa: E402
from cliora.net import cliora as ref_cliora          # noqa: E402
from cliora.net import trainer as ref_trainer        # noqa: E402
from cliora.net.inside_index import get_inside_index  # noqa: E402
from cliora.net.outside_index import get_outside_index  # noqa: E402
from cliora.net.offset_cache import get_offset_cache  # noqa: E402
from cliora.analysis.cky import ParsePredictor        # noqa: E402
from cliora.analysis.utils import override_init_with_batch, override_inside_hook  # noqa: E402


class MaskQueueDropout(torch.nn.Module):
    """Stands in for nn.Dropout(0.1) so the reference consumes explicit keep-masks."""

    def __init__(self, keep, n, p=0.1):
        super().__init__()
        self.keep, self.n, self.p, self.cursor = keep, n, p, 0

    def forward(self, x):
        L = x.shape[1]
        m = self.keep[:, self.cursor:self.cursor + L]
        self.cursor += L
        return x * m.to(x.dtype) / (1.0 - self.p)


def cotangents(g, B, C, D):
    return dict(g_inside_h=torch.randn(B, C, D, generator=g), g_inside_s=torch.randn(B, C, 1, generator=g),
                g_outside_h=torch.randn(B, C, D, generator=g), g_outside_s=torch.randn(B, C, 1, generator=g))


def named_grads(model):
    return {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
            for k, p in model.named_parameters()}


def save(name, blob):
    path = os.path.join(HERE, name)
    torch.save(blob, path)
    print('wrote', name, '%.1f KB' % (os.path.getsize(path) / 1024))


def index_fixture():
    blob = {}
    for n in range(2, 13):
        blob[('offset', n)] = get_offset_cache(n)
        for level in range(1, n):
            blob[('inside', n, level)] = get_inside_index(n, level)
        for level in range(0, n - 1):
            blob[('outside', n, level)] = get_outside_index(n, level)
    save('index.pt', blob)


def diora_case(name, B, n, D, share, seed, store_params=True):
    torch.manual_seed(seed)
    m = ref_diora.DioraMLP(D, outside=True, normalize='unit', compress=False, share=share)
    if not store_params:
        # big-D case: weights come from oracle.init_params(seed) (only the seeded draw, not the
        # algorithm) so the test can re-draw them instead of storing 3 MB of weights.
        sys.path.insert(0, os.path.join(HERE, '..', '..'))
        from oracle.cliora_oracle import init_params
        sd = init_params(D, share=share, seed=seed)
        sd['outside_compose_func.leaf_fc.weight'] = sd['inside_compose_func.leaf_fc.weight']
        sd['outside_compose_func.leaf_fc.bias'] = sd['inside_compose_func.leaf_fc.bias']
        m.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, n, D, generator=g, requires_grad=True)
    m(x, x)
    C = n * (n + 1) // 2
    ct = cotangents(g, B, C, D)
    loss = ((m.inside_h * ct['g_inside_h']).sum() + (m.inside_s * ct['g_inside_s']).sum()
            + (m.outside_h * ct['g_outside_h']).sum() + (m.outside_s * ct['g_outside_s']).sum())
    loss.backward()
    blob = dict(B=B, n=n, D=D, share=share, seed=seed, x=x.detach(),
                inside_h=m.inside_h.detach(), inside_s=m.inside_s.detach(),
                outside_h=m.outside_h.detach(), outside_s=m.outside_s.detach(),
                grad_x=x.grad.clone(), grads=named_grads(m), **ct)
    if store_params:
        blob['params'] = {k: v.detach().clone() for k, v in m.state_dict().items()}
    else:
        # keep a strided sample of the 2-D grads to bound fixture size.
        blob['grads'] = {k: (v if v.dim() == 1 else v[::16, ::16].clone()) for k, v in blob['grads'].items()}
        blob['grads_strided'] = 16
    save(name, blob)


def cliora_case(name, B, n, D, R, seed, train, E=24, V=50, K=7):
    torch.manual_seed(seed)
    m = ref_cliora.DioraMLP(D, outside=True, normalize='unit', compress=False, share=True)
    g = torch.Generator().manual_seed(seed + 1)
    C = n * (n + 1) // 2
    x_span = torch.randn(B, n, D, generator=g, requires_grad=True)
    x_word = torch.randn(B, n, D, generator=g, requires_grad=True)
    obj_span = (0.3 * torch.randn(B, R, D, generator=g)).requires_grad_()
    obj_word = (0.3 * torch.randn(B, R, D, generator=g)).requires_grad_()
    keep = (torch.rand(B, C, R, generator=g) >= 0.1)
    if train:
        m.train()
        m.atten_head.dropout = MaskQueueDropout(keep, n)
    else:
        m.eval()
    m(x_span, x_word, obj_span, obj_word)

    # losses through the reference's own loss modules (trainer.py:25-171)
    emb = torch.nn.Embedding(V, E)
    emb.weight.data = torch.randn(V, E, generator=g)
    emb.weight.requires_grad = False
    recon = ref_trainer.ReconstructionSoftmaxLoss(emb, input_size=E, size=D, k_neg=K)
    recon.mat.data = torch.randn(D, E, generator=g)
    sentences = torch.randint(0, V, (B, n), generator=g)
    neg = torch.randperm(V, generator=g)[:K]
    l_rec, _ = recon(sentences, neg, m, {})
    l_vg, _ = ref_trainer.VGLoss(alpha_vg=0.7)(sentences, m.vg_atten_score)
    l_con, _ = ref_trainer.ContrastiveLoss(margin=0.2, alpha_contr=0.9)(sentences, m)
    total = l_rec + l_vg + l_con
    total.backward()
    blob = dict(B=B, n=n, D=D, R=R, seed=seed, train=train, E=E, V=V, K=K,
                x_span=x_span.detach(), x_word=x_word.detach(), obj_span=obj_span.detach(),
                obj_word=obj_word.detach(), keep=keep,
                params={k: v.detach().clone() for k, v in m.state_dict().items()},
                inside_h=m.inside_h.detach(), inside_s=m.inside_s.detach(),
                outside_h=m.outside_h.detach(), outside_s=m.outside_s.detach(),
                all_atten_score=m.all_atten_score.detach(), vg_atten_score=m.vg_atten_score.detach(),
                atten_score=m.atten_score.detach(),
                emb_weight=emb.weight.detach(), recon_mat=recon.mat.detach().clone(),
                sentences=sentences, neg_samples=neg, alpha_vg=0.7, alpha_contr=0.9, margin=0.2,
                loss_recon=l_rec.detach(), loss_vg=l_vg.detach(), loss_contr=l_con.detach(),
                grads=named_grads(m), grad_recon_mat=recon.mat.grad.clone(),
                grad_x_span=x_span.grad.clone(), grad_x_word=x_word.grad.clone(),
                grad_obj_span=obj_span.grad.clone(), grad_obj_word=obj_word.grad.clone())
    save(name, blob)


def cky_case(name, B, n, D, seed):
    torch.manual_seed(seed)
    m = ref_diora.DioraMLP(D, outside=True, normalize='unit', compress=False, share=True)
    override_init_with_batch(m)
    override_inside_hook(m)
    m.eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, n, D, generator=g)
    with torch.no_grad():
        m(x, x)
    raw = {}
    # re-run without the max-subtracting hook to keep the RAW split scores too
    m2 = ref_diora.DioraMLP(D)
    m2.load_state_dict(m.state_dict())
    store = {}
    m2.inside_hook = types.MethodType(lambda self, level, h, c, s: store.__setitem__(level, s.detach().clone()), m2)
    with torch.no_grad():
        m2(x, x)
    trees = ParsePredictor(m).parse_batch({'sentences': torch.zeros(B, n, dtype=torch.int64)})
    save(name, dict(B=B, n=n, D=D, seed=seed, x=x, params={k: v.clone() for k, v in m.state_dict().items()},
                    split_scores=store, trees=trees))


if __name__ == '__main__':
    index_fixture()
    diora_case('diora_b2_n5_d16_share.pt', 2, 5, 16, True, 11)
    diora_case('diora_b3_n7_d32_noshare.pt', 3, 7, 32, False, 12)
    diora_case('diora_b2_n2_d16_share.pt', 2, 2, 16, True, 13)     # shortest chart: one split
    diora_case('diora_b1_n1_d16_share.pt', 1, 1, 16, True, 14)     # degenerate: single word
    diora_case('diora_b2_n6_d400_share.pt', 2, 6, 400, True, 15, store_params=False)
    cliora_case('cliora_b3_n6_d32_r5_eval.pt', 3, 6, 32, 5, 21, train=False)
    cliora_case('cliora_b3_n6_d32_r5_train.pt', 3, 6, 32, 5, 22, train=True)
    cliora_case('cliora_b4_n9_d48_r36_train.pt', 4, 9, 48, 36, 23, train=True)
    cky_case('cky_b6_n9_d24.pt', 6, 9, 24, 31)
