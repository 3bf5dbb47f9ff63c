"""Host data path around the hot path: length-bucketed sampling with the
reference's random streams, a GPU-resident region-feature table, and a prefetching batch iterator that yields
the reference's ``batch_map`` dictionaries."""
from .batch_iterator import BatchIterator, shard_indices  # noqa: F401
from .feature_store import RegionFeatureStore  # noqa: F401
from .sampler import FixedLengthBatchSampler, NegativeSampler, calculate_freq_dist  # noqa: F401
