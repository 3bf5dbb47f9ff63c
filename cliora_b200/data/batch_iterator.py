"""Batch iterator with the reference's ``batch_map`` contract (cliora/data/batch_iterator.py:44-200), built for
a GPU-resident feature table.

Differences in mechanism, not in results: the rank's shard of a batch is cut from the *index list* before
anything is loaded (the reference loads the whole global batch on every rank and chunks afterwards,
batch_iterator.py:134-136 - same ``torch.chunk`` boundaries), the visual tensors come from
``RegionFeatureStore.gather`` instead of DataLoader workers, and the next batch is assembled on a side stream
while the current step runs.
"""
import numpy as np
import torch

from .sampler import FixedLengthBatchSampler


def get_default_config():
    return dict(batch_size=16, forever=False, drop_last=False, sort_by_length=True, shuffle=True, random_seed=None,
                filter_length=None, workers=0, pin_memory=True, include_partial=False, cuda=True, ngpus=1, k_neg=3,
                negative_sampler=None, options_path=None, weights_path=None, vocab=None, length_to_size=None,
                rank=None, data_type=None, use_obj=False, mode=None, prefetch=True)


def shard_indices(index, ngpus, rank):
    """The slice of a batch that ``torch.chunk(tensor, ngpus)[rank]`` would keep (batch_iterator.py:52-66)."""
    if ngpus <= 1:
        return list(index)
    keep = torch.chunk(torch.arange(len(index)), ngpus, dim=0)
    return [index[int(i)] for i in keep[rank]] if rank < len(keep) else []


class BatchIterator(object):
    """sentences: list of token-id lists; extra: dict of per-example lists (``example_ids``, ``GT``, ``VG_GT``,
    ``image_feats`` ...) passed through into every batch_map; store: a RegionFeatureStore (or None for text-only
    data) with ``image_index[i]`` = table entry of example i (the reference's ``imgid2idx[img_ids[i]]``)."""

    def __init__(self, sentences, extra=None, store=None, image_index=None, device='cuda', **kwargs):
        self.sentences = sentences
        self.extra = extra or {}
        self.config = get_default_config()
        self.config.update({k: v for k, v in kwargs.items() if k in self.config})
        self.store = store
        self.device = torch.device(device)
        if store is not None:
            if image_index is None:
                raise ValueError('image_index is required with a feature store')
            self.image_index = np.asarray(image_index, dtype=np.int64)
            if len(self.image_index) != len(sentences):
                raise ValueError('image_index must have one entry per sentence')
        self.sampler = None
        self._stream = None

    def get_dataset_size(self):
        return len(self.sentences)

    def get_dataset_minlen(self):
        return min(map(len, self.sentences))

    def get_dataset_maxlen(self):
        return max(map(len, self.sentences))

    def get_dataset_stats(self):
        return 'size={} minlen={} maxlen={}'.format(self.get_dataset_size(), self.get_dataset_minlen(),
                                                    self.get_dataset_maxlen())

    # -- one batch, host side: ids -> pinned tensors (no device work) --
    def _host_batch(self, index, cfg):
        index = shard_indices(index, cfg['ngpus'], cfg['rank'] or 0)
        sents = torch.from_numpy(np.asarray([self.sentences[i] for i in index], dtype=np.int64))
        neg = None
        if cfg['negative_sampler'] is not None:
            neg = torch.from_numpy(cfg['negative_sampler'].sample(cfg['k_neg']))
        return index, sents, neg

    # -- device side: H2D of the ids + the gather kernel, on the current stream --
    def _device_batch(self, host):
        index, sents, neg = host
        dev = self.device
        batch_map = {'sentences': sents.pin_memory().to(dev, non_blocking=True),
                     'neg_samples': None if neg is None else neg.pin_memory().to(dev, non_blocking=True),
                     'batch_size': sents.shape[0], 'length': sents.shape[1] if sents.dim() == 2 else 0,
                     'index': tuple(index)}
        if self.store is not None:
            obj, boxes, cates = self.store.gather(self.image_index[np.asarray(index, dtype=np.int64)])
            batch_map.update(obj_feats=obj, boxes=boxes, obj_cates=cates)
        else:   # SimpleDataset placeholders (dataloader.py:116-126)
            z = torch.zeros(len(index), 1, device=dev, dtype=torch.float64)
            batch_map.update(obj_feats=z, boxes=z.clone(), obj_cates=z.long())
        for k, v in self.extra.items():
            batch_map[k] = [v[i] for i in index]
        if 'image_feats' in batch_map:
            batch_map['image_feats'] = torch.from_numpy(np.array(batch_map['image_feats']))
        return batch_map

    def get_iterator(self, **kwargs):
        cfg = dict(self.config)
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        if self.sampler is None:     # like the reference, the sampler (and its random stream) lives across epochs
            rng = np.random.RandomState(seed=cfg['random_seed'])
            self.sampler = FixedLengthBatchSampler(self.sentences, batch_size=cfg['batch_size'], rng=rng,
                                                   maxlen=cfg['filter_length'],
                                                   include_partial=cfg['include_partial'],
                                                   length_to_size=cfg['length_to_size'])
        if not cfg['prefetch']:
            return (self._device_batch(self._host_batch(ix, cfg)) for ix in self.sampler)
        return self._prefetching(cfg)

    def _prefetching(self, cfg):
        """One batch of lookahead: batch k+1 is uploaded / gathered on a side stream while the caller runs step k."""
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)
        side = self._stream

        def stage(ix):
            host = self._host_batch(ix, cfg)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                bm = self._device_batch(host)
            ev = torch.cuda.Event()
            ev.record(side)
            return bm, ev

        pending = None
        for ix in self.sampler:
            nxt = stage(ix)
            if pending is not None:
                yield self._release(pending)
            pending = nxt
        if pending is not None:
            yield self._release(pending)

    def _release(self, staged):
        bm, ev = staged
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for v in bm.values():
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(cur)
        return bm
