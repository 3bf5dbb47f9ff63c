"""GPU measurement of the batch-assembly path: gather kernel GB/s for an HBM-resident and a pinned-host table
(fp32 / fp16), and sentences/s of the whole BatchIterator (sampler + id upload + gather) with nothing consuming
the batches.  Prints one JSON object (committed as profiles/r1_data_path.json)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200.data import BatchIterator, NegativeSampler, RegionFeatureStore

B, R, F, IMAGES = 32, 36, 2048, 4000
rng = np.random.RandomState(0)
counts = rng.randint(10, 60, size=IMAGES)
ends = np.cumsum(counts)
pos = np.stack([ends - counts, ends], 1)
rows = int(ends[-1])
feats = torch.rand(rows, F)
boxes = np.random.rand(rows, 4).astype(np.float32)
out = {'config': {'batch': B, 'regions': R, 'feat_dim': F, 'images': IMAGES, 'table_rows': rows}}
for placement in ('hbm', 'pinned'):
    for dt in (torch.float32, torch.float16):
        store = RegionFeatureStore(feats, boxes, pos, regions=R, placement=placement, dtype=dt)
        idx = [torch.from_numpy(rng.randint(0, IMAGES, size=B)).cuda() for _ in range(16)]
        bufs = (torch.empty(B, R, F, device='cuda'), torch.empty(B, R, 4, device='cuda'),
                torch.empty(B, R, dtype=torch.int64, device='cuda'))
        for i in range(5):
            store.gather(idx[i], out=bufs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for i in range(reps):
            store.gather(idx[i % 16], out=bufs)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        live = float(np.minimum(counts, R).mean()) / R
        byts = B * R * F * (4 + live * (4 if dt == torch.float32 else 2))
        out['gather_%s_%s' % (placement, 'f32' if dt == torch.float32 else 'f16')] = {
            'us_per_batch': us, 'GBps': byts / us / 1e3, 'table_MB': store.table_bytes / 1e6}
        del store
store = RegionFeatureStore(feats, boxes, pos, regions=R, placement='hbm')
sents = [rng.randint(0, 8000, size=20).tolist() for _ in range(32 * 400)]
ns = NegativeSampler(np.ones(8000, dtype=np.float32), 0.75); ns.set_seed(0)
it = BatchIterator(sents, store=store, image_index=rng.randint(0, IMAGES, size=len(sents)), batch_size=B,
                   random_seed=1, negative_sampler=ns, k_neg=100)
for mode in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter(); nb = 0
    for bm in it.get_iterator(prefetch=mode):
        nb += 1
    torch.cuda.synchronize(); dt_ = time.perf_counter() - t0
    out['iterator_prefetch' if mode else 'iterator_plain'] = {'batches': nb, 'sentences_per_s': nb * B / dt_,
                                                              'ms_per_batch': dt_ / nb * 1e3}
print(json.dumps(out))
