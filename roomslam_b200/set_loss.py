"""Batched GPU Hungarian matcher + set-prediction loss: mirrors ``HungarianMatcher`` / ``SetCriterion`` of
src/benchmark/train.py:14-187 (same constructor arguments, ``criterion(outputs, targets)`` -> dict with
class_loss / l1_loss / giou_loss / total_loss) without the per-sample host round trips (train.py:36,56-57).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .functional import _need_cuda, _p, _stream


@_lib.on_tensor_device
def hungarian_match(pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid_mask, cost_class=1.0, cost_box=5.0):
    """-> (match_pred [B,K], match_slot [B,K], match_rank [B,K], n_match [B]) int32 CUDA tensors, K = min(Q, M), -1 padded."""
    _need_cuda(pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid_mask)
    B, Q = pred_boxes.shape[:2]
    M = gt_boxes.shape[1]
    K = min(Q, M)
    dev = pred_boxes.device
    pb, pc = pred_boxes.detach().contiguous().float(), pred_classes.detach().contiguous().float()
    gb, gl = gt_boxes.contiguous().float(), gt_labels.contiguous().long()
    gv = gt_valid_mask.to(torch.uint8).contiguous()
    mp = torch.empty(B, K, dtype=torch.int32, device=dev)
    ms, mr = torch.empty_like(mp), torch.empty_like(mp)
    n = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.call("rs_hungarian_match", _p(pb), _p(pc), _p(gb), _p(gl), _p(gv), B, Q, M, float(cost_class), float(cost_box),
              _p(mp), _p(ms), _p(mr), _p(n), _stream(pb))
    return mp, ms, mr, n


class HungarianMatcher:
    """train.py:14-61.  ``forward`` returns the reference's list of (pred_idx, gt_idx) arrays (one host copy for the
    whole batch); ``match`` keeps everything on the device."""

    def __init__(self, cost_class: float = 1.0, cost_box: float = 5.0):
        self.cost_class, self.cost_box = cost_class, cost_box

    def match(self, pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid_mask):
        return hungarian_match(pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid_mask, self.cost_class, self.cost_box)

    @torch.no_grad()
    def forward(self, pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid_mask) -> List[Tuple[np.ndarray, np.ndarray]]:
        mp, _, mr, n = self.match(pred_boxes, pred_classes, gt_boxes, gt_labels, gt_valid_mask)
        mp, mr, n = mp.cpu().numpy(), mr.cpu().numpy(), n.cpu().numpy()
        return [(mp[b, : n[b]].astype(np.int64), mr[b, : n[b]].astype(np.int64)) for b in range(len(n))]

    __call__ = forward


class SetLossFn(torch.autograd.Function):
    """(pred_boxes, pred_classes) + targets + matches -> losses [4] = (class, l1, giou, weighted total)."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, pred_boxes, pred_classes, gt_boxes, gt_labels, mp, ms, n, weights):
        B, Q = pred_boxes.shape[:2]
        M = gt_boxes.shape[1]
        dev = pred_boxes.device
        pb, pc = pred_boxes.contiguous().float(), pred_classes.contiguous().float()
        ws = torch.empty(1 + 3 * B, device=dev)
        losses = torch.empty(4, device=dev)
        g_logits = torch.empty(B, Q, 4, device=dev)
        g_l1, g_giou = torch.empty(B, Q, 6, device=dev), torch.empty(B, Q, 6, device=dev)
        _lib.call("rs_set_loss_f32", _p(pb), _p(pc), _p(gt_boxes), _p(gt_labels), B, Q, M, _p(mp), _p(ms), _p(n),
                  float(weights[0]), float(weights[1]), float(weights[2]), _p(ws), _p(losses), _p(g_logits), _p(g_l1),
                  _p(g_giou), _stream(pb))
        ctx.save_for_backward(g_logits, g_l1, g_giou)
        ctx.weights = weights
        return losses

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, d_losses):
        g_logits, g_l1, g_giou = ctx.saved_tensors
        wc, wl, wg = ctx.weights
        d = d_losses.float()
        d_boxes = (d[1] + wl * d[3]) * g_l1 + (d[2] + wg * d[3]) * g_giou
        d_logits = (d[0] + wc * d[3]) * g_logits
        return d_boxes, d_logits, None, None, None, None, None, None


class SetCriterion(nn.Module):
    """train.py:64-187.  ``weight_dict`` keys: class_loss, l1_loss, giou_loss (missing keys weigh 1.0, train.py:132)."""

    def __init__(self, weight_dict: Dict[str, float]):
        super().__init__()
        self.weight_dict = weight_dict
        self.matcher = HungarianMatcher()

    def forward(self, outputs, targets) -> Dict[str, torch.Tensor]:
        pb, pc = outputs["pred_boxes"], outputs["pred_classes"]
        gb = targets["boxes"].contiguous().float()
        gl = targets["labels"].contiguous().long()
        mp, ms, _, n = self.matcher.match(pb, pc, gb, gl, targets["valid_mask"])
        w = tuple(float(self.weight_dict.get(k, 1.0)) for k in ("class_loss", "l1_loss", "giou_loss"))
        losses = SetLossFn.apply(pb, pc, gb, gl, mp, ms, n, w)
        return {"class_loss": losses[0], "l1_loss": losses[1], "giou_loss": losses[2], "total_loss": losses[3]}
