"""The other BASELINE.json configurations as parity cases (bench.py only times config[1]).

c3: CKY parse inference, batch 256, length 30 -- trees vs the oracle at a batch the CPU finishes in seconds,
    plus full-size properties.
c4: long-sentence chart stress, length 64, hidden 400, batch 16 -- chart vs the oracle at batch 2, plus
    full-size properties (unit norms, sentence independence, root rules).
"""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _text_model(D=400, seed=3):
    from oracle import cliora_oracle as O
    from cliora_b200.net.diora import DioraMLP
    from test_gpu_chart import _fill
    P = O.init_params(D, seed=seed)
    m = DioraMLP(D).cuda()
    _fill(m, P)
    for k in list(P):
        if k.startswith('inside_'):
            P['outside_' + k[len('inside_'):]] = P[k]
    return m, P


def test_c3_parse_trees_length30():
    from oracle import cliora_oracle as O
    from cliora_b200.analysis.cky import ParsePredictor, backpointers
    B, n, D = 48, 30, 400
    m, P = _text_model(D)
    m.eval()
    m.outside = False                       # run_eval: text-only DIORA skips the outside pass (train.py:130)
    x = torch.randn(B, n, D, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        m(x.cuda(), x.cuda())
    trees = ParsePredictor(m).parse_batch({'sentences': torch.zeros(B, n, dtype=torch.int64)})
    out = O.chart_forward(P, x, outside=False)
    ref_best, ref_bp = O.cky_backpointers(out.split_scores, B, n)
    bp, best = backpointers(m)
    # trees identical, excluding exact-score near-ties (north_star).  A sentence may differ only if the GPU's
    # tree is a near-tie under the ORACLE's scores: its total Viterbi score (sum over its internal nodes of the
    # max-normalised split scores, cky.py:86) is within 1e-5 of the oracle's best.
    ref_trees = [O.tree_from_backpointers(ref_bp[b].tolist(), n) for b in range(B)]
    off = O.level_offsets(n)
    norm = {l: (s.reshape(B, n - l, l) - s.reshape(B, n - l, l).max(2, keepdim=True)[0])
            for l, s in out.split_scores.items()}

    def tree_score(b, row):
        def rec(level, pos):
            if level == 0:
                return 1.0
            k = int(row[off[level] + pos])
            return rec(k, pos) + rec(level - 1 - k, pos + k + 1) + float(norm[level][b, pos, k])
        return rec(n - 1, 0)
    bp_host = bp.cpu()
    for b, (t, r) in enumerate(zip(trees, ref_trees)):
        if t != r:
            top = float(ref_best[b, off[n - 1]])
            assert abs(tree_score(b, bp_host[b].tolist()) - top) <= 1e-5 * max(1.0, abs(top)), (b, t, r)
    assert rel_err(best, ref_best) < 1e-4


def test_c3_batch256_cky_kernel_bit_exact_on_its_own_scores():
    """Full c3 batch (256 x 30 words): the CKY kernel against the oracle's CKY fed with the SAME split scores
    (the ones the inside kernels produced): integer work, so every backpointer must be identical."""
    from oracle import cliora_oracle as O
    from cliora_b200.analysis.cky import backpointers
    B, n, D = 256, 30, 400
    m, _ = _text_model(D)
    m.eval()
    m.outside = False
    x = torch.randn(B, n, D, generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        m(x, x)
    bp, best = backpointers(m)
    scores = {level: m._run.split_s(level).cpu() for level in range(1, n)}
    ref_best, ref_bp = O.cky_backpointers(scores, B, n)
    assert torch.equal(bp.cpu(), ref_bp)
    assert rel_err(best, ref_best) < 1e-6


def test_c3_full_batch256_properties():
    """Full c3 size on the GPU: every tree is a valid binary bracketing of 30 leaves, decoding is
    per-sentence (a sentence parsed alone gives the same tree)."""
    from cliora_b200.analysis.cky import ParsePredictor
    B, n, D = 256, 30, 400
    m, _ = _text_model(D)
    m.eval()
    m.outside = False
    x = torch.randn(B, n, D, generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        m(x, x)
    trees = ParsePredictor(m).parse_batch({'sentences': torch.zeros(B, n, dtype=torch.int64)})

    def leaves(t):
        return [t] if isinstance(t, int) else leaves(t[0]) + leaves(t[1])
    assert all(leaves(t) == list(range(n)) for t in trees)
    with torch.no_grad():
        m(x[17:18].contiguous(), x[17:18].contiguous())
    alone = ParsePredictor(m).parse_batch({'sentences': torch.zeros(1, n, dtype=torch.int64)})
    assert alone[0] == trees[17]


def test_c4_long_sentence_chart_vs_oracle():
    from oracle import cliora_oracle as O
    B, n, D = 2, 64, 400
    m, P = _text_model(D, seed=9)
    x = torch.randn(B, n, D, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        m(x.cuda(), x.cuda())
    out = O.chart_forward(P, x)
    for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s'):
        assert rel_err(getattr(m, k), getattr(out, k)) < 1e-4, k


def test_c4_long_sentence_gradients_vs_oracle():
    """n=64 gradient parity (B=2, D=400, text-only): every chart tensor and gradient against the float64 oracle."""
    from test_gpu_chart import test_chart_vs_oracle_live
    test_chart_vs_oracle_live(2, 64, 400, 0, True)


def test_c4_full_size_properties_and_backward():
    """n=64, batch 16 (43 680 inside + 87 360 outside split rows per sentence): unit norms, root rules,
    sentence independence, finite gradients."""
    B, n, D = 16, 64, 400
    m, _ = _text_model(D, seed=9)
    C = n * (n + 1) // 2
    x = torch.randn(B, n, D, generator=torch.Generator().manual_seed(8)).cuda().requires_grad_()
    m(x, x)
    ih, oh = m.inside_h.detach(), m.outside_h.detach()
    assert (ih.norm(dim=-1) - 1).abs().max() < 1e-4 and (oh.norm(dim=-1) - 1).abs().max() < 1e-4
    root = m.root_vector_out_h.detach()
    assert rel_err(oh[:, C - 1], (root / root.norm()).expand(B, D)) < 1e-6      # diora.py:337-356
    assert m.outside_s[:, C - 1].abs().max().item() == 0 and m.inside_s[:, :n].abs().max().item() == 0
    keep_ih, keep_os = ih[5].clone(), m.outside_s.detach()[5].clone()
    (m.outside_h[:, :n].sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().max() > 0
    g = m.inside_compose_func.h_fcs[2].weight.grad
    assert torch.isfinite(g).all() and g.abs().max() > 0
    with torch.no_grad():
        m(x[5:6].detach().contiguous(), None)
    assert rel_err(m.inside_h[0], keep_ih) < 1e-5 and rel_err(m.outside_s[0], keep_os) < 1e-5
