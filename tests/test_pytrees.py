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
