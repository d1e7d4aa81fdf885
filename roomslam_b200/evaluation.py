"""Batched evaluation on the GPU: mirrors ``evaluate_metrics`` (src/benchmark/train.py:234-328) and
``post_process_predictions`` / ``nms_3d`` (src/benchmark/inference.py:87-197), and supplies the mAP the README names
(README.md:127-132) -- without the per-box-pair ``.item()`` host syncs of the shipped code.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import _lib
from .functional import _need_cuda, _p, _stream
from .set_loss import hungarian_match

LABEL_NAMES = {0: "BLOCK", 1: "LOW", 2: "MID", 3: "HIGH"}          # inference.py:135


class MetricAccumulator:
    """Device-resident counters of evaluate_metrics; ``update`` per batch (no host sync), ``compute`` once at the end."""

    def __init__(self, device="cuda", iou_thresh: float = 0.5):
        self.counts = torch.zeros(7, dtype=torch.float64, device=device)
        self.iou_thresh = float(iou_thresh)

    @_lib.on_tensor_device
    def update(self, outputs: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor]) -> None:
        pb = outputs["pred_boxes"].detach().contiguous().float()
        pc = outputs["pred_classes"].detach().contiguous().float()
        gb = targets["boxes"].contiguous().float()
        gl = targets["labels"].contiguous().long()
        gv = targets["valid_mask"].to(torch.uint8).contiguous()
        _need_cuda(pb, pc, gb, gl, gv)
        B, Q = pb.shape[:2]
        M = gb.shape[1]
        mp, ms, _, n = hungarian_match(pb, pc, gb, gl, gv)
        ws = torch.empty(7 * max(B, 1), dtype=torch.float64, device=pb.device)
        _lib.call("rs_eval_pairs", _p(pb), _p(pc), _p(gb), _p(gl), _p(gv), B, Q, M, _p(mp), _p(ms), _p(n), self.iou_thresh,
                  _p(ws), _p(self.counts), _stream(pb))

    def compute(self) -> Dict[str, float]:
        iou_sum, iou_cnt, tp, fp, fn, ok, tot = self.counts.cpu().tolist()           # train.py:313-328
        precision = tp / (tp + fp + 1e-8)
        recall = tp / (tp + fn + 1e-8)
        return {"mIoU": iou_sum / iou_cnt if iou_cnt > 0 else 0.0, "precision": precision, "recall": recall,
                "f1": 2 * precision * recall / (precision + recall + 1e-8), "cls_acc": ok / tot if tot > 0 else 0.0,
                "tp": int(tp), "fp": int(fp), "fn": int(fn)}


@torch.no_grad()
def evaluate_metrics(model, dataloader, device, iou_thresh: float = 0.5) -> Dict[str, float]:
    """Signature and result keys of train.py:234-328."""
    model.eval()
    acc = MetricAccumulator(device, iou_thresh)
    for batch in dataloader:
        outputs = model(batch["traces"].to(device), batch["trace_mask"].to(device))
        acc.update(outputs, {k: batch[k].to(device) for k in ("boxes", "labels", "valid_mask")})
    return acc.compute()


@_lib.on_tensor_device
def nms_batch(boxes: torch.Tensor, classes: torch.Tensor, confidence_threshold: float = 0.7, nms_threshold: float = 0.3):
    """boxes [B,Q,6], classes [B,Q,4] logits -> (keep_idx [B,Q] int32 (-1 padded, reference output order), n_keep [B],
    confidence [B,Q], label [B,Q])."""
    _need_cuda(boxes, classes)
    pb, pc = boxes.detach().contiguous().float(), classes.detach().contiguous().float()
    B, Q = pb.shape[:2]
    keep = torch.empty(B, Q, dtype=torch.int32, device=pb.device)
    n = torch.empty(B, dtype=torch.int32, device=pb.device)
    conf = torch.empty(B, Q, device=pb.device)
    label = torch.empty(B, Q, dtype=torch.int32, device=pb.device)
    _lib.call("rs_nms_3d", _p(pb), _p(pc), B, Q, float(confidence_threshold), float(nms_threshold), _p(keep), _p(n), _p(conf),
              _p(label), _stream(pb))
    return keep, n, conf, label


def post_process_predictions(boxes: torch.Tensor, classes: torch.Tensor, confidence_threshold: float = 0.7,
                             nms_threshold: float = 0.3) -> List[dict]:
    """Single scene ([Q,6], [Q,4]) -> the reference's list of collider dicts (inference.py:130-197)."""
    keep, n, conf, label = nms_batch(boxes[None], classes[None], confidence_threshold, nms_threshold)
    k = int(n[0])
    idx = keep[0, :k].long().cpu()
    b, c, l = boxes.detach().float().cpu(), conf[0].cpu(), label[0].cpu()
    return [{"type": "BoxCollider", "label": LABEL_NAMES[int(l[i])], "confidence": float(c[i]),
             "center": {"x": float(b[i, 0]), "y": float(b[i, 1]), "z": float(b[i, 2])},
             "size": {"x": float(b[i, 3]), "y": float(b[i, 4]), "z": float(b[i, 5])}, "radius": 0.0, "height": 0.0}
            for i in idx.tolist()]


@_lib.on_tensor_device
def ap_flags(boxes, classes, gt_boxes, gt_labels, gt_valid, iou_thresh: float = 0.5):
    """-> (confidence [B,Q], label [B,Q], tp_flag [B,Q], colliders per class [4]) on the device."""
    pb, pc = boxes.detach().contiguous().float(), classes.detach().contiguous().float()
    gb, gl = gt_boxes.contiguous().float(), gt_labels.contiguous().long()
    gv = gt_valid.to(torch.uint8).contiguous()
    _need_cuda(pb, pc, gb, gl, gv)
    B, Q = pb.shape[:2]
    flags = torch.empty(B, Q, dtype=torch.int32, device=pb.device)
    conf = torch.empty(B, Q, device=pb.device)
    label = torch.empty(B, Q, dtype=torch.int32, device=pb.device)
    n_gt = torch.zeros(4, dtype=torch.int32, device=pb.device)
    _lib.call("rs_ap_flags", _p(pb), _p(pc), _p(gb), _p(gl), _p(gv), B, Q, gb.shape[1], float(iou_thresh), _p(flags), _p(conf),
              _p(label), _p(n_gt), _stream(pb))
    return conf, label, flags, n_gt


def _average_precisions(conf, label, flags, n_gt) -> Tuple[float, List[float]]:
    """All-point interpolated AP per class from (confidence, predicted class, TP flag) of every prediction and the
    number of ground-truth objects per class; one device sort + prefix sums per class."""
    conf, label, flags = conf.flatten().double(), label.flatten(), flags.flatten().double()
    aps: List[float] = []
    for c in range(n_gt.numel()):
        total = int(n_gt[c])
        if total == 0:
            aps.append(float("nan"))
            continue
        sel = label == c
        if int(sel.sum()) == 0:
            aps.append(0.0)
            continue
        order = torch.sort(conf[sel], descending=True, stable=True).indices
        tp = flags[sel][order]
        ctp, cfp = torch.cumsum(tp, 0), torch.cumsum(1 - tp, 0)
        rec = torch.cat([ctp.new_zeros(1), ctp / total, ctp.new_ones(1)])
        prec = torch.cat([ctp.new_zeros(1), ctp / (ctp + cfp).clamp_min(1e-12), ctp.new_zeros(1)])
        prec = torch.flip(torch.cummax(torch.flip(prec, [0]), 0).values, [0])
        aps.append(float(((rec[1:] - rec[:-1]) * prec[1:]).sum()))
    valid = [a for a in aps if a == a]
    return (sum(valid) / len(valid) if valid else 0.0), aps


def mean_average_precision(boxes, classes, gt_boxes, gt_labels, gt_valid, iou_thresh: float = 0.5) -> Tuple[float, List[float]]:
    """mAP@iou_thresh over the 4 collider classes (all-point interpolated AP, classes without colliders skipped).
    The per-scene claiming runs in rs_ap_flags; the dataset-wide ranking is one device sort + prefix sums."""
    return _average_precisions(*ap_flags(boxes, classes, gt_boxes, gt_labels, gt_valid, iou_thresh))


class SlotEvaluator:
    """Evaluation of the README GRU model (``RoomSLAM``: fixed object slots + validity head, README.md:93-132) at
    BASELINE config-5 scale: ``update(pred, target)`` per chunk of traces (no host sync), ``compute()`` once.
    Metrics: mean IoU and class accuracy over the valid target slots, validity accuracy over all slots, precision /
    recall of (valid, class right, IoU >= thr) detections, and mAP@thr ranked by sigmoid(validity) * class probability."""

    def __init__(self, num_classes: int = 4, iou_thresh: float = 0.5, device="cuda"):
        self.C, self.thr = num_classes, float(iou_thresh)
        self.counts = torch.zeros(6, dtype=torch.float64, device=device)
        self.n_gt = torch.zeros(num_classes, dtype=torch.int32, device=device)
        self.n_slots = 0
        self._conf, self._label, self._flag = [], [], []

    @_lib.on_tensor_device
    def update(self, pred: Dict[str, torch.Tensor], target: Dict[str, torch.Tensor]) -> None:
        cl = pred["class_logits"].detach().contiguous().float()
        B, N, C = cl.shape
        args = [cl, pred["positions"].detach().contiguous().float(), pred["sizes"].detach().contiguous().float(),
                pred["validity_logits"].detach().contiguous().float(), target["classes"].contiguous().long(),
                target["positions"].contiguous().float(), target["sizes"].contiguous().float(), target["valid"].contiguous().float()]
        _need_cuda(*args)
        dev = cl.device
        conf = torch.empty(B, N, device=dev)
        label = torch.empty(B, N, dtype=torch.int32, device=dev)
        flag = torch.empty(B, N, dtype=torch.int32, device=dev)
        ws = torch.empty(6 * 1184, dtype=torch.float64, device=dev)
        _lib.call("rs_slot_eval", *[_p(a) for a in args], B * N, C, self.thr, _p(conf), _p(label), _p(flag), _p(self.n_gt), _p(ws),
                  _p(self.counts), _stream(cl))
        self._conf.append(conf); self._label.append(label); self._flag.append(flag)
        self.n_slots += B * N

    def compute(self) -> Dict[str, float]:
        iou_sum, n_valid, cls_hits, val_hits, tp, n_pred = self.counts.cpu().tolist()
        mAP, aps = _average_precisions(torch.cat(self._conf), torch.cat(self._label), torch.cat(self._flag), self.n_gt)
        return {"mean_iou": iou_sum / max(n_valid, 1.0), "class_accuracy": cls_hits / max(n_valid, 1.0),
                "validity_accuracy": val_hits / max(self.n_slots, 1), "precision": tp / max(n_pred, 1.0),
                "recall": tp / max(n_valid, 1.0), "mAP": mAP, "AP_per_class": aps, "n_slots": self.n_slots}
