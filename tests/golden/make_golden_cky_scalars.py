"""Golden trees from the reference's own ParsePredictor.batched_cky on a hand-made score dict (build container):

    python tests/golden/make_golden_cky_scalars.py       # writes tests/golden/cky_scalars.pt
"""
import os
import sys
import types

import torch

sys.path.insert(0, '/root/reference')
from cliora.analysis.cky import ParsePredictor  # noqa: E402


def main():
    g = torch.Generator().manual_seed(5)
    cases = []
    for B, n in [(3, 2), (4, 5), (6, 9), (2, 17)]:
        # scores as the inside hook leaves them: per-cell maximum subtracted (analysis/utils.py:86)
        raw = {level: {pos: torch.randn(B, level, generator=g) for pos in range(n - level)} for level in range(1, n)}
        scalars = {l: {p: t - t.max(1, keepdim=True)[0] for p, t in d.items()} for l, d in raw.items()}
        scalars[0] = {}
        net = types.SimpleNamespace(device=torch.device('cpu'), saved_scalars=scalars)
        pp = ParsePredictor(net)
        trees = pp.parse_batch({'sentences': torch.zeros(B, n, dtype=torch.int64)})
        cases.append(dict(B=B, n=n, scalars={l: {p: t.clone() for p, t in d.items()} for l, d in scalars.items() if l > 0},
                          trees=trees))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cky_scalars.pt')
    torch.save(cases, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
