"""Golden vectors for the tree -> actions -> spans -> stats helpers, from the reference's own functions
(cliora/analysis/utils.py:3-64), build container only:

    python tests/golden/make_golden_spans.py        # writes tests/golden/spans.json
"""
import json
import os
import random
import sys

sys.path.insert(0, '/root/reference')
from cliora.analysis.utils import get_actions, get_spans, get_stats  # noqa: E402


def random_tree(lo, hi, rng):
    if lo == hi:
        return lo
    k = rng.randint(lo, hi - 1)
    return (random_tree(lo, k, rng), random_tree(k + 1, hi, rng))


def main():
    rng = random.Random(7)
    trees = []
    for n in [2, 2, 3, 3, 4, 5, 6, 7, 9, 12, 20, 30]:
        for _ in range(3):
            t = random_tree(0, n - 1, rng)
            actions = get_actions(str(t))
            trees.append(dict(n=n, tree=str(t), actions=actions, spans=[list(s) for s in get_spans(actions)]))
    stats = []
    for _ in range(40):
        a = {(rng.randint(0, 6), rng.randint(0, 6)) for _ in range(rng.randint(0, 8))}
        b = {(rng.randint(0, 6), rng.randint(0, 6)) for _ in range(rng.randint(0, 8))}
        stats.append(dict(a=sorted(a), b=sorted(b), stats=list(get_stats(a, b))))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'spans.json')
    json.dump(dict(trees=trees, stats=stats), open(path, 'w'))
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
