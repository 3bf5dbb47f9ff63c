"""Region-feature table for the visual half of a CLIORA batch, resident on the GPU side.

The reference keeps the MAF features in host numpy arrays and builds every batch on DataLoader workers
(``FlickrDataset``, cliora/data/dataloader.py:188-225), then copies ``[B, 36, 2048]`` fp32 to the device per
step (cliora/data/batch_iterator.py:163-166).  Here the whole ragged table is placed once either in HBM
(Flickr30K train is 8.8 GB in fp32, 4.4 GB in fp16 - small against 180 GB) or in pinned host memory that the
gather kernel reads zero-copy, and a batch is assembled by one kernel launch (``cliora_gather_regions``).
"""
import numpy as np
import torch

from .. import _lib
from .._lib import ClioraError, check, ptr


def _table_ptr(t):
    """Address of a table tensor: device memory, or page-locked host memory (unified addressing makes the same
    pointer valid inside a kernel)."""
    if t is None:
        return None
    if t.is_cuda:
        return ptr(t)
    if not (t.is_pinned() and t.is_contiguous()):
        raise ClioraError('cliora_b200: host-resident table tensors must be pinned and contiguous')
    return t.data_ptr()


class RegionFeatureStore(object):
    """features [rows, F] (fp32 or fp16), bboxes [rows, 4], pos_bboxes [images, 2] (start, end rows), optional
    classes [rows] - the arrays ``FlickrDataset`` reads from its HDF5 file (dataloader.py:200-203).

    placement='hbm': the table lives on ``device``; 'pinned': in page-locked host memory, read by the kernel
    directly.  ``dtype=torch.float16`` stores the features in half precision (values are widened to fp32 when a
    batch is gathered; the step itself stays fp32)."""

    def __init__(self, features, bboxes, pos_bboxes, classes=None, regions=36, device='cuda', placement='hbm',
                 dtype=None):
        if placement not in ('hbm', 'pinned'):
            raise ValueError("placement must be 'hbm' or 'pinned'")
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('cliora_b200: RegionFeatureStore needs a CUDA device (no CPU path)')
        feats = torch.as_tensor(np.asarray(features) if not torch.is_tensor(features) else features)
        if dtype is not None:
            feats = feats.to(dtype)
        if feats.dtype not in (torch.float32, torch.float16):
            feats = feats.float()
        if feats.dim() != 2 or feats.shape[1] % 8:
            raise RuntimeError('cliora_b200: features must be [rows, F] with F % 8 == 0')
        pos = torch.as_tensor(np.asarray(pos_bboxes)).to(torch.int64).reshape(-1, 2)
        if pos.numel() and (int(pos.min()) < 0 or int(pos[:, 1].max()) > feats.shape[0]
                            or bool((pos[:, 1] < pos[:, 0]).any())):
            raise RuntimeError('cliora_b200: pos_bboxes outside the feature table')
        box = torch.as_tensor(np.asarray(bboxes)).float().reshape(-1, 4)
        if box.shape[0] != feats.shape[0]:
            raise RuntimeError('cliora_b200: bboxes and features disagree on the row count')
        cls = None if classes is None else torch.as_tensor(np.asarray(classes)).to(torch.int32).reshape(-1)
        place = (lambda t: t.contiguous().to(self.device)) if placement == 'hbm' else \
            (lambda t: t.contiguous().pin_memory())
        self.features, self.bboxes, self.pos = place(feats), place(box), place(pos)
        self.classes = None if cls is None else place(cls)
        self.placement = placement
        self.regions = int(regions)
        self.num_images = pos.shape[0]
        self.F = feats.shape[1]

    @property
    def table_bytes(self):
        return self.features.numel() * self.features.element_size()

    def gather(self, img_index, out=None):
        """img_index: int64 ids into pos_bboxes (host list/array/tensor or device tensor).
        Returns (obj_feats [B,R,F] f32, boxes [B,R,4] f32, obj_cates [B,R] i64) on the device, filled on the
        current stream."""
        if torch.is_tensor(img_index) and img_index.is_cuda:
            idx = img_index.to(torch.int64).contiguous()      # caller vouches for the range
        else:
            host = torch.as_tensor(np.asarray(img_index)).to(torch.int64).reshape(-1)
            if host.numel() and (int(host.min()) < 0 or int(host.max()) >= self.num_images):
                raise IndexError('cliora_b200: image index out of range')
            idx = host.pin_memory().to(self.device, non_blocking=True)
        B, R, F = idx.numel(), self.regions, self.F
        if out is None:
            out = (torch.empty(B, R, F, device=self.device), torch.empty(B, R, 4, device=self.device),
                   torch.empty(B, R, dtype=torch.int64, device=self.device))
        obj, boxes, cates = out
        if B:
            check(_lib.lib().cliora_gather_regions(
                B, R, F, 1 if self.features.dtype == torch.float16 else 0, _table_ptr(self.features),
                _table_ptr(self.bboxes), _table_ptr(self.classes), _table_ptr(self.pos), ptr(idx), ptr(obj), ptr(boxes),
                ptr(cates), _lib.stream()), 'cliora_gather_regions')
        return obj, boxes, cates
