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
