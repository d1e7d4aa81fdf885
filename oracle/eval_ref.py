"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline may import this; the product path
never does).  numpy / torch-CPU restatement of the upstream evaluation code:

  * iou_pairs, batch_counts, metrics_from_counts <- evaluate_metrics   src/benchmark/train.py:234-328
  * nms_order                                    <- nms_3d + post_process_predictions  src/benchmark/inference.py:87-170
  * average_precision                            <- README.md:127-132 promises mAP; the repository ships no code for it.
        Defined here as: per class, predictions of all scenes sorted by confidence; a prediction is a true positive if
        its best-IoU not-yet-claimed collider of the same class in the same scene has IoU >= 0.5 (axis-aligned boxes,
        as in all shipped IoU code); AP = area under the all-point interpolated precision/recall curve; mAP = mean over
        the classes that have colliders.  PARITY UNPINNED for this one function (spec only).

PARITY PINNED for the rest: tests/golden/eval.npz holds outputs of the reference's own evaluate_metrics and
post_process_predictions (oracle/make_golden_eval.py).
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import set_loss_ref


def iou_pairs(pb: torch.Tensor, gb: torch.Tensor) -> torch.Tensor:
    """train.py:277-292 (note: eps is added to the union BEFORE the division, unlike the loss)."""
    lo = torch.maximum(pb[:, :3] - pb[:, 3:] / 2, gb[:, :3] - gb[:, 3:] / 2)
    hi = torch.minimum(pb[:, :3] + pb[:, 3:] / 2, gb[:, :3] + gb[:, 3:] / 2)
    inter = (hi - lo).clamp(min=0).prod(1)
    return inter / (pb[:, 3:].prod(1) + gb[:, 3:].prod(1) - inter + 1e-6)


def batch_counts(outputs, targets, iou_thresh=0.5):
    """-> dict(iou_sum, iou_cnt, tp, fp, fn, cls_correct, cls_total) for one batch (train.py:247-311)."""
    boxes, logits = outputs["pred_boxes"], outputs["pred_classes"]
    labels_pred = logits.softmax(-1).argmax(-1)
    pairs = set_loss_ref.match(boxes, logits, targets["boxes"], targets["labels"], targets["valid_mask"])
    c = dict(iou_sum=0.0, iou_cnt=0, tp=0, fp=0, fn=0, cls_correct=0, cls_total=0)
    for b, (pi, gi) in enumerate(pairs):
        keep = targets["valid_mask"][b]
        c["fn"] += max(0, int(keep.sum()) - len(gi))
        if len(pi) == 0:
            continue
        ious = iou_pairs(boxes[b, pi], targets["boxes"][b, keep][gi])
        c["iou_sum"] += float(ious.sum()); c["iou_cnt"] += ious.numel()
        c["cls_correct"] += int((labels_pred[b, pi] == targets["labels"][b, keep][gi]).sum()); c["cls_total"] += len(pi)
        c["tp"] += int((ious >= iou_thresh).sum()); c["fp"] += int((ious < iou_thresh).sum())
    return c


def metrics_from_counts(c):
    """train.py:313-328."""
    precision = c["tp"] / (c["tp"] + c["fp"] + 1e-8)
    recall = c["tp"] / (c["tp"] + c["fn"] + 1e-8)
    return {"mIoU": c["iou_sum"] / c["iou_cnt"] if c["iou_cnt"] else 0.0, "precision": precision, "recall": recall,
            "f1": 2 * precision * recall / (precision + recall + 1e-8),
            "cls_acc": c["cls_correct"] / c["cls_total"] if c["cls_total"] else 0.0, "tp": c["tp"], "fp": c["fp"], "fn": c["fn"]}


def iou_one(a: np.ndarray, b: np.ndarray) -> float:
    """inference.py:60-84 in fp32."""
    a, b = a.astype(np.float32), b.astype(np.float32)
    lo = np.maximum(a[:3] - a[3:] / 2, b[:3] - b[3:] / 2)
    hi = np.minimum(a[:3] + a[3:] / 2, b[:3] + b[3:] / 2)
    inter = np.prod(np.clip(hi - lo, 0, None), dtype=np.float32)
    return float(inter / (np.prod(a[3:], dtype=np.float32) + np.prod(b[3:], dtype=np.float32) - inter + np.float32(1e-6)))


def nms_order(boxes: torch.Tensor, logits: torch.Tensor, conf_thr=0.7, nms_thr=0.3):
    """Query indices kept by post_process_predictions, in its output order (class 0..3, descending confidence within a
    class), with their labels and confidences (inference.py:130-170)."""
    probs = torch.softmax(logits, -1)
    conf, label = probs.max(-1)
    out = []
    for cls in range(4):
        idx = [int(i) for i in torch.nonzero((conf > conf_thr) & (label == cls)).flatten()]
        idx.sort(key=lambda i: -float(conf[i]))
        alive = idx
        while alive:                                              # inference.py:106-125
            cur, rest = alive[0], alive[1:]
            out.append(cur)
            alive = [i for i in rest if iou_one(boxes[cur].numpy(), boxes[i].numpy()) < nms_thr]
    return np.array(out, np.int64), label.numpy(), conf.numpy()


def average_precision(scores, tp_flags, n_gt):
    if n_gt == 0:
        return float("nan")
    order = np.argsort(-np.asarray(scores, np.float64), kind="stable")
    tp = np.asarray(tp_flags, np.float64)[order]
    ctp, cfp = np.cumsum(tp), np.cumsum(1 - tp)
    rec = np.concatenate([[0.0], ctp / n_gt, [1.0]])
    prec = np.concatenate([[0.0], ctp / np.maximum(ctp + cfp, 1e-12), [0.0]])
    for i in range(len(prec) - 2, -1, -1):
        prec[i] = max(prec[i], prec[i + 1])
    step = np.nonzero(rec[1:] != rec[:-1])[0]
    return float(np.sum((rec[step + 1] - rec[step]) * prec[step + 1]))


def map_flags(boxes, logits, gt_boxes, gt_labels, gt_valid, iou_thr=0.5):
    """Per prediction: (confidence, predicted class, TP flag) with greedy claiming by descending confidence within
    (scene, class); plus the number of colliders per class."""
    probs = torch.softmax(logits, -1)
    conf, label = probs.max(-1)
    B, Q = conf.shape
    flags = np.zeros((B, Q), np.int32)
    n_gt = np.zeros(4, np.int64)
    for b in range(B):
        slots = [int(m) for m in torch.nonzero(gt_valid[b]).flatten()]
        for m in slots:
            n_gt[int(gt_labels[b, m])] += 1
        claimed = set()
        for q in sorted(range(Q), key=lambda q: -float(conf[b, q])):
            best, best_m = -1.0, -1
            for m in slots:
                if m in claimed or int(gt_labels[b, m]) != int(label[b, q]):
                    continue
                v = iou_one(boxes[b, q].numpy(), gt_boxes[b, m].numpy())
                if v > best:
                    best, best_m = v, m
            if best_m >= 0 and best >= iou_thr:
                claimed.add(best_m)
                flags[b, q] = 1
    return conf.numpy(), label.numpy(), flags, n_gt


def mean_average_precision(boxes, logits, gt_boxes, gt_labels, gt_valid, iou_thr=0.5):
    conf, label, flags, n_gt = map_flags(boxes, logits, gt_boxes, gt_labels, gt_valid, iou_thr)
    aps = [average_precision(conf[label == c], flags[label == c], int(n_gt[c])) for c in range(4)]
    valid = [a for a in aps if not np.isnan(a)]
    return (float(np.mean(valid)) if valid else 0.0), aps


def slot_eval(pred, target, iou_thr=0.5):
    """Evaluation of the README GRU model's slot outputs (spec only, README.md:93-132: PARITY UNPINNED): the torch-CPU
    definition of what roomslam_b200.evaluation.SlotEvaluator computes."""
    cl, pos, size, vl = pred["class_logits"], pred["positions"], pred["sizes"], pred["validity_logits"]
    tv = target["valid"] > 0.5
    label = cl.argmax(-1)
    conf = torch.sigmoid(vl) * torch.softmax(cl, -1).max(-1).values
    lo = torch.maximum(pos - size / 2, target["positions"] - target["sizes"] / 2)
    hi = torch.minimum(pos + size / 2, target["positions"] + target["sizes"] / 2)
    inter = (hi - lo).clamp_min(0).prod(-1)
    iou = inter / (size.prod(-1) + target["sizes"].prod(-1) - inter).clamp_min(1e-9)
    flag = tv & (label == target["classes"]) & (iou >= iou_thr)
    pv = torch.sigmoid(vl) > 0.5
    C = cl.shape[-1]
    n_gt = np.array([int((tv & (target["classes"] == c)).sum()) for c in range(C)])
    aps = [average_precision(conf[label == c].numpy(), flag[label == c].numpy().astype(np.float64), int(n_gt[c])) for c in range(C)]
    ok = [a for a in aps if not np.isnan(a)]
    nv = max(int(tv.sum()), 1)
    return {"mean_iou": float(iou[tv].sum()) / nv, "class_accuracy": float((label == target["classes"])[tv].sum()) / nv,
            "validity_accuracy": float((pv == tv).float().mean()), "precision": float(flag.sum()) / max(int(pv.sum()), 1),
            "recall": float(flag.sum()) / nv, "mAP": float(np.mean(ok)) if ok else 0.0, "AP_per_class": aps}
