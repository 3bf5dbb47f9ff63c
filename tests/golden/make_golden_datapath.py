"""Golden vectors for the visual batch assembly from the reference's own FlickrDataset.__getitem__
(cliora/data/dataloader.py:205-222), build container only:

    python tests/golden/make_golden_datapath.py

h5py is absent here and only used by FlickrDataset.__init__ (which reads files we do not have), so the instance is
created without running __init__ and given the arrays __init__ would have loaded; __getitem__ itself runs
unmodified.  Writes tests/golden/datapath.npz.
"""
import os
import sys
import types

import numpy as np
sys.modules.setdefault('h5py', types.ModuleType('h5py'))
sys.path.insert(0, '/root/reference')
from cliora.data.dataloader import FlickrDataset  # noqa: E402


def main():
    rng = np.random.RandomState(3)
    images, F = 4, 2048                               # the reference hard-codes 2048-d features
    counts = np.array([0, 40, 3, 2])                  # an empty image and one with more than 36 boxes
    ends = np.cumsum(counts)
    pos = np.stack([ends - counts, ends], 1).astype(np.int64)
    rows = int(ends[-1])
    feats = rng.randint(0, 256, size=(rows, F)).astype(np.float32) / 16    # few distinct values: compresses well
    boxes = (rng.rand(rows, 4) * 500).astype(np.float32)
    classes = rng.randint(0, 40, size=rows).astype(np.int32)
    examples = 5
    img_of_example = np.array([1, 0, 2, 3, 1])

    ds = object.__new__(FlickrDataset)
    ds.dataset = [list(range(5))] * examples
    ds.img_ids = [str(1000 + int(i)) for i in img_of_example]
    ds.mode = 'train'
    ds.imgid2idx = {1000 + i: i for i in range(images)}
    ds.obj2ind = {'c%d' % i: i for i in range(40)}
    ds.detection_dict = {str(1000 + i): {'classes': ['c%d' % c for c in classes[pos[i, 0]:pos[i, 1]]]}
                         for i in range(images)}
    ds.features, ds.predicted_boxes, ds.indexes = feats, boxes, pos
    items = [ds[i] for i in range(examples)]
    out = dict(features=feats, bboxes=boxes, pos=pos, classes=classes, image_index=img_of_example.astype(np.int64),
               obj_feats=np.array([it[2] for it in items]), boxes=np.array([it[3] for it in items]),
               obj_cates=np.array([it[4] for it in items]).astype(np.int64))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'datapath.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
