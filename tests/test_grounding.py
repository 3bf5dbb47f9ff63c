"""Phrase-grounding scoring: oracle pinned against torchvision (CPU), device kernel against the oracle (GPU)."""
import pytest
import torch

from oracle import grounding as og


def _case(B, n, R, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    atten = torch.randn(B, n, R, generator=g)
    if ties:   # quantise so that equal maxima occur and the first-max rule matters
        atten = (atten * 2).round() / 2
    xy = torch.rand(B, R, 2, generator=g) * 300
    wh = torch.rand(B, R, 2, generator=g) * 200 + 1
    boxes = torch.cat([xy, xy + wh], -1)
    targets = []
    for b in range(B):
        t = {}
        for j in range(int(torch.randint(0, 4, (1,), generator=g))):
            s = int(torch.randint(0, n, (1,), generator=g))
            e = int(torch.randint(s + 1, n + 1, (1,), generator=g))
            r = int(torch.randint(0, R, (1,), generator=g))
            jit = (torch.rand(4, generator=g) - 0.5) * 80       # some boxes land on either side of IoU 0.5
            t['p%d' % j] = (s, e, (boxes[b, r] + jit).tolist())
        targets.append((t, None))
    return atten, boxes, targets


def test_oracle_iou_matches_torchvision():
    tv = pytest.importorskip('torchvision.ops')
    g = torch.Generator().manual_seed(0)
    for _ in range(200):
        a = torch.rand(4, generator=g) * 100
        b = torch.rand(4, generator=g) * 100
        a[2:] += a[:2]
        b[2:] += b[:2]
        assert float(og.box_iou(a, b)) == float(tv.box_iou(a[None], b[None])[0, 0])


@pytest.mark.parametrize('ties', [False, True])
def test_oracle_two_spellings_agree(ties):
    atten, boxes, targets = _case(6, 12, 36, 3, ties)
    assert og.ground_phrases(atten, boxes, targets) == og.ground_phrases_train_py(atten, boxes, targets)


@pytest.mark.gpu
@pytest.mark.parametrize('B,n,R,seed,ties', [(8, 20, 36, 1, False), (8, 20, 36, 2, True), (3, 5, 70, 4, True),
                                             (16, 30, 36, 5, False)])
def test_grounding_kernel_vs_oracle(B, n, R, seed, ties):
    from cliora_b200.analysis.grounding import flatten_targets, grounding_eval, grounding_recall
    atten, boxes, targets = _case(B, n, R, seed, ties)
    want = og.ground_phrases(atten, boxes, targets)
    phrases, gts = flatten_targets(targets)
    sel, iou, hit = grounding_eval(atten.cuda(), boxes.cuda(), phrases, gts)
    assert len(want) == phrases.shape[0]
    assert sel.cpu().tolist() == [[w[3], w[4]] for w in want]
    assert iou.cpu().tolist() == pytest.approx([w[5] for w in want], rel=0, abs=0)     # same fp32 arithmetic
    assert hit.cpu().tolist() == [w[6] for w in want]

    class _D:
        atten_score = atten.cuda()
    rec, tot, res = grounding_recall(_D, {'VG_GT': targets, 'boxes': boxes})
    assert (rec, tot) == (sum(w[6] for w in want), len(want))
    assert [len(r) for r in res] == [len(t[0]) for t in targets]


@pytest.mark.gpu
def test_grounding_no_phrases():
    from cliora_b200.analysis.grounding import flatten_targets, grounding_eval
    phrases, gts = flatten_targets([({}, None), ({}, None)])
    sel, iou, hit = grounding_eval(torch.randn(2, 4, 5).cuda(), torch.rand(2, 5, 4).cuda(), phrases, gts)
    assert sel.shape == (0, 2) and iou.numel() == 0 and hit.numel() == 0


def test_oracle_matches_reference_code_golden():
    """tests/golden/grounding.pt was produced by executing the reference's own scoring blocks (cut out of
    scripts/parse.py and scripts/train.py, see make_golden_grounding.py): the oracle must reproduce them."""
    import os
    cases = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'grounding.pt'), weights_only=False)
    assert sum(c['recall_num'] for c in cases) > 0 and sum(c['total_num'] - c['recall_num'] for c in cases) > 0
    for c in cases:
        got = og.ground_phrases(c['atten'], c['boxes'], c['targets'])
        per_sentence = [[] for _ in c['targets']]
        for bid, s, e, _, _, _, hit in got:
            per_sentence[bid].append(((s, e - 1), hit))
        assert per_sentence == c['ground_res']
        assert sum(g[6] for g in got) == c['recall_num'] and len(got) == c['total_num']


@pytest.mark.gpu
def test_grounding_kernel_vs_reference_code_golden():
    import os
    from cliora_b200.analysis.grounding import grounding_recall
    cases = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'grounding.pt'), weights_only=False)
    for c in cases:
        class _D:
            atten_score = c['atten'].cuda()
        rec, tot, res = grounding_recall(_D, {'VG_GT': c['targets'], 'boxes': c['boxes']})
        assert (rec, tot) == (c['recall_num'], c['total_num'])
        assert res == c['ground_res']
