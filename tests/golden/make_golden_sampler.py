"""Golden batch orders from the UNMODIFIED reference sampler / negative sampler (run in the build container only):

    python tests/golden/make_golden_sampler.py

cliora/data/dataloader.py imports h5py at module level (absent here, and only used by FlickrDataset's
constructor), so a stub module is registered before the import; the sampler code itself runs untouched.
Writes tests/golden/sampler.json.
"""
import json
import os
import sys
import types

import numpy as np

sys.modules.setdefault('h5py', types.ModuleType('h5py'))
sys.path.insert(0, '/root/reference')
from cliora.blocks.negative_sampler import NegativeSampler  # noqa: E402
from cliora.data.dataloader import FixedLengthBatchSampler, SimpleDataset  # noqa: E402


def sentences(seed, count, lo, hi):
    rng = np.random.RandomState(seed)
    return [list(range(int(l))) for l in rng.randint(lo, hi + 1, size=count)]


CASES = [
    dict(name='plain', seed=1, count=300, lo=3, hi=12, batch_size=8, rng_seed=11),
    dict(name='partial', seed=2, count=257, lo=2, hi=9, batch_size=16, rng_seed=5, include_partial=True),
    dict(name='maxlen', seed=3, count=200, lo=3, hi=30, batch_size=4, rng_seed=7, maxlen=20),
    dict(name='length_to_size', seed=4, count=400, lo=3, hi=25, batch_size=32, rng_seed=9, include_partial=True,
         length_to_size={10: 16, 20: 8}),
    dict(name='two_epochs', seed=5, count=120, lo=4, hi=8, batch_size=8, rng_seed=3, epochs=2),
]


def main():
    out = {'sampler': [], 'negative': []}
    for c in CASES:
        sents = sentences(c['seed'], c['count'], c['lo'], c['hi'])
        ds = SimpleDataset(sents)
        l2s = c.get('length_to_size')
        s = FixedLengthBatchSampler(ds, batch_size=c['batch_size'], rng=np.random.RandomState(seed=c['rng_seed']),
                                    maxlen=c.get('maxlen'), include_partial=c.get('include_partial', False),
                                    length_to_size=l2s)
        epochs = [[list(map(int, b)) for b in s] for _ in range(c.get('epochs', 1))]
        rec = dict(c)
        if l2s:
            rec['length_to_size'] = {str(k): v for k, v in l2s.items()}
        rec['lengths'] = [len(x) for x in sents]
        rec['epochs'] = epochs
        out['sampler'].append(rec)
    for seed, V, power, k in [(0, 50, 0.75, 10), (7, 1000, 0.75, 100), (3, 200, 0.0, 5)]:
        freq = np.random.RandomState(100 + seed).randint(0, 40, size=V).astype(np.float32)
        ns = NegativeSampler(freq, power)
        ns.set_seed(seed)
        draws = [ns.sample(k).tolist() for _ in range(3)]
        out['negative'].append(dict(seed=seed, V=V, power=power, k=k, freq=freq.tolist(), draws=draws))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sampler.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
