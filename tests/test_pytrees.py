"""Host helper csrc/pytrees.c (nested-tuple trees from a backpointer table) against the Python recursion and the
oracle's follow-backpointers restatement."""
import random

import pytest
import torch

from cliora_b200 import _lib
from cliora_b200.analysis.cky import tree_from_backpointers, trees_from_table


def _table(B, n, seed):
    rng = random.Random(seed)
    C = n * (n + 1) // 2
    off = [l * n - l * (l - 1) // 2 for l in range(n)]
    rows = []
    for _ in range(B):
        r = [-1] * C
        for l in range(1, n):
            for p in range(n - l):
                r[off[l] + p] = rng.randrange(l)
        rows.append(r)
    return torch.tensor(rows, dtype=torch.int32).reshape(B, C)


@pytest.mark.parametrize('B,n', [(1, 1), (3, 2), (7, 9), (64, 30), (2, 64)])
def test_c_helper_matches_python_and_oracle(B, n):
    from oracle.cliora_oracle import tree_from_backpointers as oracle_tree
    _lib.build_pytrees()
    from cliora_b200 import _pytrees
    bp = _table(B, n, seed=n)
    got = _pytrees.build(bp.numpy(), B, n)
    assert got == [tree_from_backpointers(r, n) for r in bp.tolist()]
    assert got == [oracle_tree(bp[b], n) for b in range(B)]
    assert trees_from_table(bp, n) == got


def test_c_helper_rejects_bad_input():
    _lib.build_pytrees()
    from cliora_b200 import _pytrees
    with pytest.raises(ValueError):
        _pytrees.build(torch.full((1, 3), 7, dtype=torch.int32).numpy(), 1, 2)     # backpointer out of range
    with pytest.raises(ValueError):
        _pytrees.build(torch.zeros(2, dtype=torch.int32).numpy(), 1, 3)            # buffer too small


def _cky_scalar_cases():
    import os
    return torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'cky_scalars.pt'), weights_only=False)


def test_pack_scalars_layout_against_reference_batched_cky():
    """tests/golden/cky_scalars.pt: trees the reference's own ParsePredictor produced from a score dict.  Packing
    that dict into the kernel's flat layout and decoding with the oracle's CKY must give the same trees - this pins
    the layout batched_cky() hands to the kernel (the kernel itself is bit-exact against the oracle on the GPU)."""
    from cliora_b200.analysis.cky import pack_scalars
    from oracle import cliora_oracle as O
    for c in _cky_scalar_cases():
        B, n = c['B'], c['n']
        flat = pack_scalars(c['scalars'], B, n)
        scores, o = {}, 0
        for level in range(1, n):
            rows = B * (n - level) * level
            scores[level] = flat[o:o + rows].reshape(B, n - level, level, 1)
            o += rows
        assert o == flat.numel()
        _, bp = O.cky_backpointers(scores, B, n)
        assert [O.tree_from_backpointers(bp[b].tolist(), n) for b in range(B)] == c['trees']


@pytest.mark.gpu
def test_batched_cky_from_scalars_vs_reference_golden():
    from cliora_b200.analysis.cky import ParsePredictor
    import types
    for c in _cky_scalar_cases():
        net = types.SimpleNamespace(device=torch.device('cuda', 0))
        scalars = {l: {p: t.cuda() for p, t in d.items()} for l, d in c['scalars'].items()}
        trees = ParsePredictor(net).batched_cky({'sentences': torch.zeros(c['B'], c['n'], dtype=torch.int64)}, scalars)
        assert trees == c['trees']
