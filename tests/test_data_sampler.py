"""Host data path: the sampler / negative sampler mirrors reproduce the reference's batches exactly
(tests/golden/sampler.json, made by tests/golden/make_golden_sampler.py from the reference classes)."""
import json
import os

import numpy as np
import pytest

from cliora_b200.data.sampler import FixedLengthBatchSampler, NegativeSampler, calculate_freq_dist

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'sampler.json')))


@pytest.mark.parametrize('case', GOLD['sampler'], ids=[c['name'] for c in GOLD['sampler']])
def test_sampler_matches_reference(case):
    sents = [list(range(l)) for l in case['lengths']]
    l2s = {int(k): v for k, v in case['length_to_size'].items()} if case.get('length_to_size') else None
    s = FixedLengthBatchSampler(sents, batch_size=case['batch_size'], rng=np.random.RandomState(case['rng_seed']),
                                maxlen=case.get('maxlen'), include_partial=case.get('include_partial', False),
                                length_to_size=l2s)
    for want in case['epochs']:
        got = [list(b) for b in s]
        assert got == want
        assert len(s) == len(want)
        for b in got:                                  # the property the chart kernels rely on
            assert len({case['lengths'][i] for i in b}) == 1


def test_sampler_accepts_length_array_and_data_source():
    lengths = np.array([3, 5, 3, 5, 5, 3, 4, 4], dtype=np.int64)

    class Src:
        dataset = [list(range(l)) for l in lengths]

        def __len__(self):
            return len(self.dataset)
    a = list(FixedLengthBatchSampler(lengths, 2, rng=np.random.RandomState(1), include_partial=True))
    b = list(FixedLengthBatchSampler(Src(), 2, rng=np.random.RandomState(1), include_partial=True))
    assert a == b and sorted(i for batch in a for i in batch) == list(range(8))


@pytest.mark.parametrize('case', GOLD['negative'], ids=lambda c: 'V%d' % c['V'])
def test_negative_sampler_matches_reference(case):
    ns = NegativeSampler(np.asarray(case['freq'], dtype=np.float32), case['power'])
    ns.set_seed(case['seed'])
    assert [ns.sample(case['k']).tolist() for _ in range(3)] == case['draws']


def test_freq_dist():
    f = calculate_freq_dist([[0, 1, 1], [3, 1]], 5)
    assert f.tolist() == [1, 3, 0, 1, 0] and f.dtype == np.float32


@pytest.mark.parametrize('size,ngpus', [(32, 2), (32, 8), (7, 2), (9, 4), (5, 1)])
def test_shard_before_loading_equals_chunk_after_loading(size, ngpus):
    """The rank's slice of the index list == what the reference keeps after torch.chunk on the loaded batch
    (batch_iterator.py:52-66,134-136), for every rank."""
    import torch
    from cliora_b200.data import shard_indices
    index = [100 + i for i in range(size)]
    loaded = torch.arange(size) + 100
    pieces = torch.chunk(loaded, ngpus, dim=0)
    for rank in range(len(pieces)):
        assert shard_indices(index, ngpus, rank) == pieces[rank].tolist()
    got = [i for rank in range(ngpus) for i in shard_indices(index, ngpus, rank)]
    assert got == index          # a partition of the batch, in order


def test_feature_store_has_no_cpu_path():
    import numpy as np
    from cliora_b200.data import RegionFeatureStore
    with pytest.raises(RuntimeError):
        RegionFeatureStore(np.zeros((4, 8), np.float32), np.zeros((4, 4), np.float32), np.array([[0, 4]]),
                           device='cpu')


def test_datapath_oracle_matches_reference_getitem():
    """tests/golden/datapath.npz: outputs of the reference's own FlickrDataset.__getitem__
    (tests/golden/make_golden_datapath.py); the oracle's restatement must reproduce them exactly."""
    import os
    import torch
    from oracle import data_path as od
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'datapath.npz'))
    items = [od.flickr_item(z['features'], z['bboxes'], z['pos'], z['classes'], int(i)) for i in z['image_index']]
    want = od.collate([[0]] * 64, list(range(len(items))), items)
    assert torch.equal(want['obj_feats'], torch.from_numpy(z['obj_feats']))
    assert torch.equal(want['boxes'], torch.from_numpy(z['boxes']))
    assert torch.equal(want['obj_cates'], torch.from_numpy(z['obj_cates']))


@pytest.mark.skipif(not os.path.isdir('/root/reference/cliora'), reason='needs the reference checkout')
def test_sampler_matches_reference_live_on_random_configurations():
    """Beyond the committed golden cases: 40 random (data set, batch size, maxlen, length_to_size, include_partial)
    configurations against the reference class itself, two epochs each (the random stream carries over)."""
    import sys
    import types
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))
    sys.path.insert(0, '/root/reference')
    from cliora.data.dataloader import FixedLengthBatchSampler as RefSampler, SimpleDataset
    rng = np.random.RandomState(123)
    for trial in range(40):
        count = int(rng.randint(1, 300))
        lo = int(rng.randint(1, 6))
        hi = lo + int(rng.randint(0, 25))
        sents = [list(range(int(l))) for l in rng.randint(lo, hi + 1, size=count)]
        kw = dict(batch_size=int(rng.randint(1, 40)), include_partial=bool(rng.randint(0, 2)),
                  maxlen=[None, 0, int(rng.randint(lo, hi + 2))][int(rng.randint(0, 3))],
                  length_to_size=[None, {int(rng.randint(2, 15)): int(rng.randint(1, 20)),
                                         int(rng.randint(15, 30)): int(rng.randint(1, 10))}][int(rng.randint(0, 2))])
        seed = int(rng.randint(0, 1000))
        ref = RefSampler(SimpleDataset(sents), rng=np.random.RandomState(seed), **kw)
        mine = FixedLengthBatchSampler(sents, rng=np.random.RandomState(seed), **kw)
        for _ in range(2):
            assert [list(map(int, b)) for b in ref] == [list(b) for b in mine], (trial, kw)
            assert len(ref) == len(mine)
