"""CPU restatement of the reference's phrase-grounding scoring.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
nothing under cliora_b200/ may import this.

The reference has no function for it: the logic is inlined in its eval loops
(cliora/scripts/parse.py:174-212 and cliora/scripts/train.py:158-179, two spellings of the same selection).
``ground_phrases`` follows parse.py line by line on CPU tensors; ``ground_phrases_train_py`` follows train.py;
tests check the two agree and pin ``box_iou`` against torchvision.ops.box_iou (what the reference calls).
Pinning: tests/golden/grounding.pt holds the outputs of the reference's own scoring blocks, cut out of the
unmodified script files and executed on seeded inputs (tests/golden/make_golden_grounding.py);
tests/test_grounding.py checks this restatement against them, and the IoU arithmetic against torchvision.
"""
import torch


def box_iou(a, b):
    """torchvision.ops.box_iou for one box pair each ([4] tensors): intersection / union, areas (x2-x1)(y2-y1)."""
    area_a = (a[2] - a[0]) * (a[3] - a[1])
    area_b = (b[2] - b[0]) * (b[3] - b[1])
    lt = torch.max(a[:2], b[:2])
    rb = torch.min(a[2:], b[2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[0] * wh[1]
    return inter / (area_a + area_b - inter)


def ground_phrases(atten_score, boxes, targets, thresh=0.5):
    """parse.py:174-212.  atten_score [B,n,R], boxes [B,R,4] CPU; targets = VG_GT.  Returns a list of
    (bid, start, end, word, region, iou, hit) in visiting order."""
    out = []
    for bid in range(len(targets)):
        target_bid = targets[bid][0]
        for _, (start_id, end_id, gt_box) in target_bid.items():
            words_scores = atten_score[bid][start_id:end_id]
            max_word_scores, _ = words_scores.max(1)
            select_wid = int(max_word_scores.max(0)[1])
            word2phr_atten = words_scores[select_wid]
            select_box = int(word2phr_atten.max(0)[1])
            iou = box_iou(boxes[bid][select_box], torch.tensor(gt_box, dtype=torch.float32))
            out.append((bid, start_id, end_id, start_id + select_wid, select_box, float(iou), int(iou > thresh)))
    return out


def ground_phrases_train_py(atten_score, boxes, targets, thresh=0.5):
    """train.py:158-179 (per-word argmax first, then the best word of the phrase)."""
    out = []
    for bid in range(len(targets)):
        target_bid = targets[bid][0]
        select_scores, select_box_ids = atten_score[bid].max(1)
        pred_boxes = boxes[bid][select_box_ids]
        for _, (start_id, end_id, gt_box) in target_bid.items():
            select_id = int(select_scores[start_id:end_id].max(0)[1])
            iou = box_iou(pred_boxes[start_id:end_id][select_id], torch.tensor(gt_box, dtype=torch.float32))
            out.append((bid, start_id, end_id, start_id + select_id, int(select_box_ids[start_id + select_id]),
                        float(iou), int(iou > thresh)))
    return out
