"""Golden vectors for the phrase-grounding scoring, produced by the reference's OWN code (build container only):

    python tests/golden/make_golden_grounding.py

The scoring is not a function in the reference: it is a block inside the eval loops of
cliora/scripts/parse.py (and, spelled differently, cliora/scripts/train.py).  This script cuts those blocks out of
the unmodified source files by their first and last statements, dedents them and executes them on seeded inputs
(``diora.atten_score``, ``batch_map['boxes']``, ``batch_map['VG_GT']``); nothing is re-typed.  Writes
tests/golden/grounding.pt.
"""
import os
import textwrap
import types

import torch
import torchvision.ops as torchops

REF = '/root/reference/cliora/scripts/'


def cut(path, first, last):
    lines = open(path).read().split('\n')
    i0 = next(i for i, l in enumerate(lines) if l.strip() == first)
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip() == last)
    return textwrap.dedent('\n'.join(lines[i0:i1 + 1]))


PARSE_BLOCK = cut(REF + 'parse.py', 'batch_ground_res = None', 'batch_ground_res.append(ground_res)')
TRAIN_BLOCK = cut(REF + 'train.py', 'if diora.atten_score is not None:', 'total_num += 1')


def case(B, n, R, seed, ties):
    g = torch.Generator().manual_seed(seed)
    atten = torch.randn(B, n, R, generator=g)
    if ties:
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
            jit = (torch.rand(4, generator=g) - 0.5) * 60
            if float(torch.rand(1, generator=g)) < 0.6:      # make the annotated region the one the phrase attends to,
                atten[b, s, r] += 6.0                        # so that both IoU outcomes occur
            t['p%d' % j] = (s, e, (boxes[b, r] + jit).tolist())
        targets.append((t, None))
    return atten, boxes, targets


def run(block, atten, boxes, targets):
    ns = {'diora': types.SimpleNamespace(atten_score=atten), 'batch_map': {'VG_GT': targets, 'boxes': boxes},
          'torch': torch, 'torchops': torchops, 'recall_num': 0, 'total_num': 0}
    exec(block, ns)
    return ns


def main():
    out = []
    for B, n, R, seed, ties in [(8, 20, 36, 1, False), (8, 20, 36, 2, True), (3, 5, 70, 4, True), (16, 30, 36, 5, False)]:
        atten, boxes, targets = case(B, n, R, seed, ties)
        p = run(PARSE_BLOCK, atten, boxes, targets)
        t = run(TRAIN_BLOCK, atten, boxes, targets)
        assert (p['recall_num'], p['total_num']) == (t['recall_num'], t['total_num'])
        out.append(dict(atten=atten, boxes=boxes, targets=targets, ground_res=p['batch_ground_res'],
                        recall_num=p['recall_num'], total_num=p['total_num']))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'grounding.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes;', [(c['recall_num'], c['total_num']) for c in out])


if __name__ == '__main__':
    main()
