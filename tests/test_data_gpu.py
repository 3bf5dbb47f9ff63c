"""Device batch assembly against the oracle's restatement of FlickrDataset.__getitem__ + collate."""
import numpy as np
import pytest
import torch

from oracle import data_path as od


def _table(seed, images=40, F=64, max_rows=50):
    rng = np.random.RandomState(seed)
    counts = rng.randint(0, max_rows + 1, size=images)       # includes empty images and images with > 36 rows
    counts[0], counts[1] = 0, max_rows
    ends = np.cumsum(counts)
    pos = np.stack([ends - counts, ends], 1).astype(np.int64)
    rows = int(ends[-1])
    feats = rng.rand(rows, F).astype(np.float32)
    boxes = (rng.rand(rows, 4) * 500).astype(np.float32)
    classes = rng.randint(0, 1600, size=rows).astype(np.int32)
    return feats, boxes, pos, classes


@pytest.mark.gpu
@pytest.mark.parametrize('placement', ['hbm', 'pinned'])
@pytest.mark.parametrize('half', [False, True])
def test_gather_regions_vs_oracle(placement, half):
    from cliora_b200.data import RegionFeatureStore
    feats, boxes, pos, classes = _table(3)
    store = RegionFeatureStore(feats, boxes, pos, classes, regions=36, placement=placement,
                               dtype=torch.float16 if half else None)
    idx = [0, 1, 7, 7, 39, 12, 1]
    obj, bx, ct = store.gather(idx)
    ref_feats = feats.astype(np.float16).astype(np.float32) if half else feats
    items = [od.flickr_item(ref_feats, boxes, pos, classes, i) for i in idx]
    want = od.collate([[0]] * 64, list(range(len(idx))), items)
    assert torch.equal(obj.cpu(), want['obj_feats'])          # a copy (and an exact fp16 -> fp32 widening)
    assert torch.equal(bx.cpu(), want['boxes'])
    assert torch.equal(ct.cpu(), want['obj_cates'])


@pytest.mark.gpu
def test_gather_rejects_bad_index():
    from cliora_b200.data import RegionFeatureStore
    feats, boxes, pos, classes = _table(4, images=5)
    store = RegionFeatureStore(feats, boxes, pos, classes)
    with pytest.raises(IndexError):
        store.gather([0, 5])


@pytest.mark.gpu
@pytest.mark.parametrize('ngpus,rank', [(1, 0), (2, 0), (2, 1)])
def test_batch_iterator_vs_oracle(ngpus, rank):
    """Same sampler stream -> same batches; shard cut before loading == reference's chunk after loading."""
    from cliora_b200.data import BatchIterator, FixedLengthBatchSampler, NegativeSampler, RegionFeatureStore
    feats, boxes, pos, classes = _table(5, images=30, F=32)
    rng = np.random.RandomState(0)
    sents = [rng.randint(0, 100, size=int(l)).tolist() for l in rng.randint(3, 8, size=200)]
    image_index = rng.randint(0, 30, size=200)
    extra = {'example_ids': ['ex%d' % i for i in range(200)], 'GT': [[(0, 1)]] * 200}
    store = RegionFeatureStore(feats, boxes, pos, classes)
    ns = NegativeSampler(np.ones(100, dtype=np.float32), 0.75)
    ns.set_seed(3)
    it = BatchIterator(sents, extra=extra, store=store, image_index=image_index, batch_size=8, random_seed=13,
                       ngpus=ngpus, rank=rank, negative_sampler=ns, k_neg=5, include_partial=ngpus == 1)
    ns2 = NegativeSampler(np.ones(100, dtype=np.float32), 0.75)
    ns2.set_seed(3)
    # (the reference cannot chunk a partial batch smaller than ngpus: torch.chunk returns too few pieces)
    ref_sampler = FixedLengthBatchSampler(sents, 8, rng=np.random.RandomState(13), include_partial=ngpus == 1)
    count = 0
    for bm, index in zip(it.get_iterator(), ref_sampler):
        items = [od.flickr_item(feats, boxes, pos, classes, image_index[i]) for i in index]
        want = od.collate(sents, index, items, ngpus, rank)
        assert bm['index'] == tuple(want['index'])
        assert torch.equal(bm['sentences'].cpu(), want['sents'])
        assert torch.equal(bm['obj_feats'].cpu(), want['obj_feats'])
        assert torch.equal(bm['boxes'].cpu(), want['boxes'])
        assert torch.equal(bm['obj_cates'].cpu(), want['obj_cates'])
        assert bm['neg_samples'].cpu().tolist() == ns2.sample(5).tolist()
        assert bm['example_ids'] == [extra['example_ids'][i] for i in want['index']]
        assert bm['batch_size'] == want['sents'].shape[0] and bm['length'] == want['sents'].shape[1]
        count += 1
    assert count == len(ref_sampler) > 10


@pytest.mark.gpu
def test_batch_iterator_feeds_trainer():
    """batch_maps from the iterator drive Trainer.step unchanged (keys of trainer.py:437-448)."""
    from test_gpu_trainer import _trainer
    from cliora_b200.data import BatchIterator, NegativeSampler, RegionFeatureStore
    rng = np.random.RandomState(1)
    D, F, R, V = 32, 2048, 6, 50
    counts = rng.randint(1, 9, size=20)
    ends = np.cumsum(counts)
    pos = np.stack([ends - counts, ends], 1)
    store = RegionFeatureStore(rng.rand(int(ends[-1]), F).astype(np.float32), rng.rand(int(ends[-1]), 4), pos,
                               regions=R)
    sents = [rng.randint(0, V, size=6).tolist() for _ in range(16)]
    ns = NegativeSampler(np.ones(V, dtype=np.float32), 0.75)
    ns.set_seed(0)
    it = BatchIterator(sents, store=store, image_index=rng.randint(0, 20, size=16), batch_size=4, random_seed=2,
                       negative_sampler=ns, k_neg=7)
    trainer = _trainer(D=D, V=V, k_neg=7)
    losses = [trainer.step(bm)['total_loss'] for bm in it.get_iterator()]
    assert len(losses) == 4 and all(np.isfinite(l) for l in losses)


@pytest.mark.gpu
def test_gather_regions_vs_reference_golden():
    """The gather kernel reproduces what the reference's own FlickrDataset.__getitem__ returned (golden)."""
    import os
    from cliora_b200.data import RegionFeatureStore
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'datapath.npz'))
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    store = RegionFeatureStore(g['features'], g['bboxes'], g['pos'], g['classes'], regions=36)
    obj, bx, ct = store.gather(g['image_index'])
    assert torch.equal(obj.cpu(), g['obj_feats'])
    assert torch.equal(bx.cpu(), g['boxes'])
    assert torch.equal(ct.cpu(), g['obj_cates'])
