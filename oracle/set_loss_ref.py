"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline may import this; the product path
never does).  torch-CPU + scipy restatement of the upstream set-prediction loss:

  * match            <- HungarianMatcher.forward   src/benchmark/train.py:21-61
  * giou_pairs       <- SetCriterion.box_iou_3d    src/benchmark/train.py:75-107
  * set_loss         <- SetCriterion.forward       src/benchmark/train.py:109-187

The assignment solver itself is the third-party scipy.optimize.linear_sum_assignment (train.py:4,57; scipy is in the
image on both boxes): a shortest-augmenting-path solver after D. F. Crouse, "On implementing 2D rectangular assignment
algorithms", IEEE TAES 52(4), 2016.

PARITY PINNED: tests/golden/set_loss.npz holds the matches, losses and gradients of the reference's own SetCriterion
(imported from /root/reference by oracle/make_golden_set_loss.py); tests/test_oracle_set_loss.py checks this file
against them.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

WEIGHTS = {"class_loss": 2.0, "l1_loss": 5.0, "giou_loss": 2.0}       # train.py:433-437
COST_CLASS, COST_BOX = 1.0, 5.0                                      # train.py:17


@torch.no_grad()
def match(pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid):
    """-> list over the batch of (pred_idx, gt_idx) int arrays; gt_idx counts within the VALID colliders of the sample."""
    pairs = []
    for b in range(pred_boxes.shape[0]):
        keep = gt_valid[b]
        if int(keep.sum()) == 0:
            pairs.append((np.zeros(0, np.int64), np.zeros(0, np.int64)))
            continue
        prob = pred_classes[b].softmax(-1)
        cost = COST_CLASS * (-prob[:, gt_labels[b, keep]]) + COST_BOX * torch.cdist(pred_boxes[b], gt_boxes[b, keep], p=1)
        rows, cols = linear_sum_assignment(cost.cpu().numpy())
        pairs.append((rows.astype(np.int64), cols.astype(np.int64)))
    return pairs


def giou_pairs(a, b):
    """Axis-aligned 3-D IoU / GIoU of paired boxes (cx, cy, cz, sx, sy, sz), eps 1e-6 in both quotients."""
    a_lo, a_hi = a[:, :3] - a[:, 3:] / 2, a[:, :3] + a[:, 3:] / 2
    b_lo, b_hi = b[:, :3] - b[:, 3:] / 2, b[:, :3] + b[:, 3:] / 2
    inter = (torch.minimum(a_hi, b_hi) - torch.maximum(a_lo, b_lo)).clamp(min=0).prod(1)
    union = a[:, 3:].prod(1) + b[:, 3:].prod(1) - inter
    iou = inter / (union + 1e-6)
    hull = (torch.maximum(a_hi, b_hi) - torch.minimum(a_lo, b_lo)).clamp(min=0).prod(1)
    return iou, iou - (hull - union) / (hull + 1e-6)


def set_loss(outputs, targets, weights=WEIGHTS, pairs=None):
    boxes, logits = outputs["pred_boxes"], outputs["pred_classes"]
    gt_boxes, gt_labels, gt_valid = targets["boxes"], targets["labels"], targets["valid_mask"]
    if pairs is None:
        pairs = match(boxes, logits, gt_boxes, gt_labels, gt_valid)
    sel_logits, sel_labels, sel_boxes, sel_gt = [], [], [], []
    for b, (pi, gi) in enumerate(pairs):
        if len(pi) == 0:
            continue
        pi_t, gi_t = torch.as_tensor(pi, device=boxes.device), torch.as_tensor(gi, device=boxes.device)
        sel_logits.append(logits[b, pi_t])
        sel_labels.append(gt_labels[b, gt_valid[b]][gi_t])
        sel_boxes.append(boxes[b, pi_t])
        sel_gt.append(gt_boxes[b, gt_valid[b]][gi_t])
    zero = torch.tensor(0.0, device=boxes.device)
    if not sel_logits:
        losses = {"class_loss": zero, "l1_loss": zero, "giou_loss": zero}
    else:
        pl, tl, pb, tb = torch.cat(sel_logits), torch.cat(sel_labels), torch.cat(sel_boxes), torch.cat(sel_gt)
        losses = {"class_loss": torch.nn.functional.cross_entropy(pl, tl),
                  "l1_loss": (pb - tb).abs().mean(),
                  "giou_loss": (1 - giou_pairs(pb, tb)[1]).mean()}
    losses["total_loss"] = sum(losses[k] * weights.get(k, 1.0) for k in ("class_loss", "l1_loss", "giou_loss"))
    return losses, pairs
