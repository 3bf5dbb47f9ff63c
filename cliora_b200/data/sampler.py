"""Length-bucketed batch sampling and negative sampling with the reference's random streams.

Mirrors ``FixedLengthBatchSampler`` (cliora/data/dataloader.py:11-113) and ``NegativeSampler``
(cliora/blocks/negative_sampler.py:28-38): the same ``np.random.RandomState`` yields the same batches in the
same order (tests/golden/sampler.json is produced by the reference classes).  The chart kernels need a fixed
sentence length per batch, which is exactly what this sampler guarantees.
"""
import numpy as np


def _lengths_of(source):
    """Sentence lengths from a reference-style data source (``.dataset[i]`` is a token list), a list of token
    lists, or an integer array of lengths."""
    items = getattr(source, 'dataset', source)
    if isinstance(items, np.ndarray) and items.ndim == 1 and np.issubdtype(items.dtype, np.integer):
        return items.astype(np.int64)
    return np.fromiter((len(x) for x in items), dtype=np.int64, count=len(items))


class FixedLengthBatchSampler(object):
    """Yields lists of example ids that all have the same sentence length.

    ``length_to_size`` maps a length threshold to the batch size used from that length on (piecewise constant,
    dataloader.py:27-38); ``include_partial`` appends each bucket's remainder as a short batch
    (dataloader.py:80-84); ``maxlen`` drops longer sentences (dataloader.py:58-59)."""

    def __init__(self, data_source, batch_size, include_partial=False, rng=None, maxlen=None, length_to_size=None):
        self.data_source = data_source
        self.lengths = _lengths_of(data_source)
        self.rng = np.random.RandomState(seed=11) if rng is None else rng
        self.batch_size = batch_size
        self.maxlen = maxlen
        self.include_partial = include_partial
        self.length_to_size = length_to_size
        self._thresholds = sorted((int(k), int(v)) for k, v in (length_to_size or {}).items())
        self.order = []
        self.length_map = {}
        self._cursor = {}
        self.index = -1

    def get_batch_size(self, length):
        size = self.batch_size
        for threshold, value in self._thresholds:
            if 0 < threshold <= length:
                size = value
        return size

    def reset(self):
        lengths = self.lengths
        keep = np.arange(len(lengths))
        if self.maxlen is not None and self.maxlen > 0:
            keep = keep[lengths[keep] <= self.maxlen]
        # buckets in order of first appearance (the reference builds a dict while scanning the data set, and the
        # shuffles below consume the random stream bucket by bucket in that order)
        uniq, first = np.unique(lengths[keep], return_index=True)
        buckets = [int(l) for l in uniq[np.argsort(first, kind='stable')]]
        self.length_map = {}
        for l in buckets:
            ids = [int(i) for i in keep[lengths[keep] == l]]
            self.rng.shuffle(ids)
            self.length_map[l] = ids
        order, tail = [], []
        for l in buckets:
            size = self.get_batch_size(l)
            full, rest = divmod(len(self.length_map[l]), size)
            order += [l] * full
            if rest and self.include_partial:
                tail.append(l)
        order += tail
        self.rng.shuffle(order)
        self.order = order
        self._cursor = {l: 0 for l in buckets}
        self.index = -1

    def get_next_batch(self):
        self.index += 1
        length = self.order[self.index]
        size = self.get_batch_size(length)
        start = self._cursor[length]
        self._cursor[length] = start + size
        return self.length_map[length][start:start + size]

    def __iter__(self):
        self.reset()
        for _ in range(len(self.order)):
            yield self.get_next_batch()

    def __len__(self):
        return len(self.order)


class NegativeSampler(object):
    """Unigram^power negative sampler without replacement (negative_sampler.py:28-38)."""

    def __init__(self, freq_dist, dist_power, epsilon=10 ** -2):
        freq_dist = np.asarray(freq_dist, dtype=np.float32)
        dist = freq_dist ** dist_power + epsilon * (1 / len(freq_dist))
        self.dist = dist / sum(dist)     # python sum: the reference's (sequential fp32) normaliser
        self.rng = np.random.RandomState()

    def set_seed(self, seed):
        self.rng.seed(seed)

    def sample(self, num_samples):
        return self.rng.choice(len(self.dist), num_samples, p=self.dist, replace=False)


def calculate_freq_dist(data, vocab_size):
    """Token counts over a list of id sequences (negative_sampler.py:15-25), vectorised."""
    counts = np.zeros(vocab_size, dtype=np.int64)
    for x in data:
        np.add.at(counts, np.asarray(x, dtype=np.int64), 1)
    return counts.astype(np.float32)
