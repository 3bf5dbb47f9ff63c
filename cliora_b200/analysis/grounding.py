"""Phrase-grounding recall on the device.

Replaces the per-phrase host loops of the reference's eval scripts (cliora/scripts/parse.py:174-212,
cliora/scripts/train.py:158-179), which copy ``diora.atten_score`` and the boxes to the CPU and call
``torchvision.ops.box_iou`` once per phrase.  Targets keep the reference's structure: ``batch_map['VG_GT']`` is a
list (one entry per sentence) of ``(target, noun_mask)`` where ``target`` maps a phrase id to
``(start_id, end_id, gt_box)`` with ``end_id`` exclusive and ``gt_box = [x1, y1, x2, y2]``.
"""
import torch

from .. import _lib
from .._lib import check, ptr


def flatten_targets(targets):
    """VG_GT -> (phrases [P,3] int32 = (sentence, start, end), gt_boxes [P,4] f32), in the reference's visiting
    order (sentences in batch order, phrases in dict order)."""
    phrases, gts = [], []
    for bid, entry in enumerate(targets):
        target = entry[0] if isinstance(entry, (tuple, list)) else entry
        for _, (start_id, end_id, gt_box) in target.items():
            phrases.append((bid, int(start_id), int(end_id)))
            gts.append([float(v) for v in gt_box])
    return (torch.tensor(phrases, dtype=torch.int32).reshape(-1, 3),
            torch.tensor(gts, dtype=torch.float32).reshape(-1, 4))


def grounding_eval(atten_score, boxes, phrases, gt_boxes, iou_thresh=0.5):
    """atten_score [B,n,R] (``diora.atten_score``), boxes [B,R,4], phrases [P,3] int32, gt_boxes [P,4], all CUDA.
    Returns (sel [P,2] int32 = selected (word, region), iou [P] f32, hit [P] int32)."""
    if not atten_score.is_cuda:
        raise RuntimeError('cliora_b200: grounding_eval needs CUDA tensors (no CPU path)')
    dev = atten_score.device
    B, n, R = atten_score.shape
    atten_score = atten_score.contiguous().float()
    boxes = boxes.to(dev, torch.float32).contiguous()
    phrases = phrases.to(dev, torch.int32).contiguous()
    gt_boxes = gt_boxes.to(dev, torch.float32).contiguous()
    if boxes.shape != (B, R, 4):
        raise RuntimeError('cliora_b200: boxes must be [B, R, 4], got %s' % (tuple(boxes.shape),))
    P = phrases.shape[0]
    sel = torch.empty(P, 2, dtype=torch.int32, device=dev)
    iou = torch.empty(P, dtype=torch.float32, device=dev)
    hit = torch.empty(P, dtype=torch.int32, device=dev)
    if P:
        if int(phrases[:, 0].min()) < 0 or int(phrases[:, 0].max()) >= B:
            raise RuntimeError('cliora_b200: phrase sentence index out of range')
        check(_lib.lib().cliora_grounding_eval(B, n, R, P, ptr(atten_score), ptr(boxes), ptr(phrases), ptr(gt_boxes),
                                               float(iou_thresh), ptr(sel), ptr(iou), ptr(hit), _lib.stream()),
              'cliora_grounding_eval')
    return sel, iou, hit


def grounding_recall(diora, batch_map, iou_thresh=0.5):
    """(recall_num, total_num, per-sentence [((start, end-1), hit), ...]) for one batch, like the accumulation
    in scripts/parse.py:174-212; ``diora.atten_score`` must be live (eval-mode forward with object features)."""
    targets = batch_map['VG_GT']
    phrases, gts = flatten_targets(targets)
    _, _, hit = grounding_eval(diora.atten_score, batch_map['boxes'], phrases, gts, iou_thresh)
    hit = hit.cpu().tolist()
    res = [[] for _ in targets]
    for (bid, s, e), h in zip(phrases.tolist(), hit):
        res[bid].append(((s, e - 1), int(h)))
    return sum(hit), len(hit), res
