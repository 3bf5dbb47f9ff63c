"""CPU restatement of the reference's visual batch assembly.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
nothing under cliora_b200/ may import this.

``flickr_item`` follows FlickrDataset.__getitem__ (cliora/data/dataloader.py:205-222) on the arrays that class
reads from its HDF5 file; ``collate`` follows collate_fn + the rank partition
(cliora/data/batch_iterator.py:116-138).  The HDF5/pickle/json lookups in front (img_id -> feat_index,
class name -> id) are replaced by already-resolved integer arrays; there is no arithmetic in them to pin.
Pinning: tests/golden/datapath.npz holds what the reference's own FlickrDataset.__getitem__ returned on a small
seeded table (tests/golden/make_golden_datapath.py); tests/test_data_sampler.py checks ``flickr_item`` against it.
"""
import numpy as np
import torch


def flickr_item(features, predicted_boxes, indexes, classes, feat_index, regions=36):
    start, end = indexes[feat_index]
    num_box = min(end - start, regions)
    boxes = np.zeros([regions, 4]).astype(np.float32) - 1
    boxes[:num_box] = predicted_boxes[start:end][:num_box]
    obj_feats = np.zeros([regions, features.shape[1]]).astype(np.float32)
    obj_feats[:num_box] = features[start:end][:num_box]
    obj_cates = np.zeros([regions]).astype(np.int32) - 1
    if classes is not None:
        obj_cates[:num_box] = classes[start:end][:num_box]
    return obj_feats, boxes, obj_cates


def collate(sentences, index, items, ngpus=1, rank=0):
    obj_feats, boxes, obj_cates = zip(*items)
    out = {'index': tuple(index),
           'sents': torch.from_numpy(np.array([sentences[i] for i in index])).long(),
           'obj_feats': torch.from_numpy(np.array(obj_feats)),
           'boxes': torch.from_numpy(np.array(boxes)),
           'obj_cates': torch.from_numpy(np.array(obj_cates)).long()}
    if ngpus > 1:
        for k, v in out.items():
            if isinstance(v, torch.Tensor):
                out[k] = torch.chunk(v, ngpus, dim=0)[rank]
            else:
                keep = torch.chunk(torch.arange(len(v)), ngpus, dim=0)[rank]
                out[k] = [v[int(i)] for i in keep]
    return out
