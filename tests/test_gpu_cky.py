"""CKY kernel: trees identical to the reference's ParsePredictor (golden) and to the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cky_trees_vs_golden(golden):
    from cliora_b200.net.diora import DioraMLP
    from cliora_b200.analysis.cky import ParsePredictor
    from cliora_b200.analysis.utils import override_init_with_batch, override_inside_hook
    from test_gpu_chart import _fill
    blob = golden('cky_b6_n9_d24.pt')
    m = DioraMLP(blob['D']).cuda()
    _fill(m, blob['params'])
    override_init_with_batch(m)      # the reference's parse.py installs both hooks (parse.py:119-120)
    override_inside_hook(m)
    m.eval()
    with torch.no_grad():
        m(blob['x'].cuda(), blob['x'].cuda())
    assert set(m.saved_scalars.keys()) == set(range(blob['n']))
    trees = ParsePredictor(m).parse_batch({'sentences': torch.zeros(blob['B'], blob['n'], dtype=torch.int64)})
    assert trees == blob['trees']


@pytest.mark.parametrize('B,n,D', [(16, 30, 400), (4, 64, 64), (3, 1, 16), (3, 2, 16)])
def test_cky_vs_oracle_live(B, n, D):
    """Bigger charts: kernel backpointers vs the oracle's CKY on the kernel's own split scores
    (bit-exact decode), and vs the oracle end-to-end wherever the decision margin is clear."""
    from oracle import cliora_oracle as O
    from cliora_b200.net.diora import DioraMLP
    from cliora_b200.analysis.cky import backpointers
    from test_gpu_chart import _fill
    m = DioraMLP(D).cuda()
    P = O.init_params(D, seed=3)
    _fill(m, P)
    x = torch.randn(B, n, D, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        m(x.cuda(), x.cuda())
    bp, best = backpointers(m)
    scores = {l: m._run.split_s(l).cpu() for l in range(1, n)}
    ref_best, ref_bp = O.cky_backpointers(scores, B, n)
    assert torch.equal(bp.cpu(), ref_bp)            # integer/index work: bit-exact on identical inputs
    assert torch.equal(best.cpu(), ref_best)
    if n <= 30:
        # end-to-end against the CPU oracle forward: identical trees, excluding near-ties
        for k in list(P):
            if k.startswith('inside_'):
                P['outside_' + k[len('inside_'):]] = P[k]
        out = O.chart_forward(P, x, outside=False)
        o_best, o_bp = O.cky_backpointers(out.split_scores, B, n)
        same = (o_bp == bp.cpu())
        if not same.all():
            # any disagreement must be a near-tie in the oracle's own candidate scores
            assert (o_best - best.cpu()).abs().max() < 1e-3
            assert same.float().mean() > 0.99


@pytest.mark.parametrize('B,n', [(7, 9), (64, 30), (3, 2)])
def test_device_spans_match_host_tree_spans(B, n):
    """tree_spans_kernel == spans read off the nested-tuple trees (reference: get_actions/get_spans on str(tree))."""
    from oracle import cliora_oracle as O
    from cliora_b200.net.diora import DioraMLP
    from cliora_b200.analysis.cky import ParsePredictor
    from cliora_b200.analysis.utils import get_spans_from_tree
    from test_gpu_chart import _fill
    m = DioraMLP(32).cuda()
    _fill(m, O.init_params(32, seed=2))
    x = torch.randn(B, n, 32, generator=torch.Generator().manual_seed(n)).cuda()
    with torch.no_grad():
        m(x, x)
    pp = ParsePredictor(m)
    trees = pp.parse_batch({'sentences': torch.zeros(B, n, dtype=torch.int64)})
    sp = pp.parse_spans().cpu().tolist()
    for b in range(B):
        assert [tuple(s) for s in sp[b]] == get_spans_from_tree(trees[b])   # same post-order, same spans
        assert sp[b][-1] == [0, n - 1]


def test_device_span_f1_matches_reference_set_arithmetic():
    """span_f1_kernel == get_stats + the sentence-F1 arithmetic of scripts/parse.py:216-233 (set semantics)."""
    import random
    from oracle import cliora_oracle as O
    from cliora_b200.net.diora import DioraMLP
    from cliora_b200.analysis.cky import ParsePredictor, span_f1
    from cliora_b200.analysis.utils import get_spans_from_tree
    from test_gpu_chart import _fill
    B, n = 12, 11
    m = DioraMLP(32).cuda()
    _fill(m, O.init_params(32, seed=4))
    x = torch.randn(B, n, 32, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        m(x, x)
    trees = ParsePredictor(m).parse_batch({'sentences': torch.zeros(B, n, dtype=torch.int64)})
    rnd = random.Random(0)
    gold = []
    for b in range(B):
        own = get_spans_from_tree(trees[b])
        g = [s for s in own if rnd.random() < 0.5]                       # some right ones
        g += [(rnd.randrange(0, n - 2), rnd.randrange(2, n)) for _ in range(3)]   # some arbitrary ones
        g += g[:1]                                                          # a duplicate (set semantics)
        g.append((0, n - 1))                                                # the last entry is dropped
        gold.append(g if b != 5 else [(0, n - 1)])                          # one sentence with an empty gold set
    got = span_f1(m, gold).cpu()
    for b in range(B):
        gs = set(tuple(s) for s in gold[b][:-1])
        ps = set(get_spans_from_tree(trees[b])[:-1])
        tp = len(ps & gs); fp = len(ps - gs); fn = len(gs - ps)
        prec = float(tp) / (len(ps) + 1e-8)
        reca = float(tp) / (len(gs) + 1e-8)
        if len(gs) == 0:
            reca = 1.
            if len(ps) == 0:
                prec = 1.
        f1 = 2 * prec * reca / (prec + reca + 1e-8)
        assert got[b, :3].tolist() == [tp, fp, fn], (b, got[b], tp, fp, fn)
        assert abs(got[b, 3].item() - f1) < 1e-5
