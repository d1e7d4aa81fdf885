"""Generates tests/golden/eval.npz from the REFERENCE's own evaluate_metrics (src/benchmark/train.py:234-328) and
post_process_predictions (src/benchmark/inference.py:130-197), imported from /root/reference.

    python -m oracle.make_golden_eval [--reference /root/reference]

evaluate_metrics is driven with a stub model that returns seeded predictions and a list of batch dicts as the
dataloader.  Predictions are made to overlap the colliders (collider + noise) so that TP / FP / NMS suppression all occur.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"trained": (30, 50, 16, 31), "wide": (80, 50, 5, 32), "narrow": (5, 50, 7, 33)}
LABELS = {"BLOCK": 0, "LOW": 1, "MID": 2, "HIGH": 3}       # inference.py:135 output label map, inverted


def case_inputs(name):
    Q, M, B, seed = CASES[name]
    g = torch.Generator().manual_seed(seed)
    gt_boxes = torch.cat([torch.randn(B, M, 3, generator=g) * 3, torch.rand(B, M, 3, generator=g) * 2 + 0.3], -1)
    labels = torch.randint(0, 4, (B, M), generator=g)
    valid = torch.rand(B, M, generator=g) < torch.rand(B, 1, generator=g)
    valid[0] = False
    # predictions: noisy copies of (cyclically chosen) valid colliders, so IoUs straddle the thresholds and duplicates exist
    boxes = torch.empty(B, Q, 6)
    logits = torch.randn(B, Q, 4, generator=g)
    for b in range(B):
        slots = torch.nonzero(valid[b]).flatten()
        for q in range(Q):
            if len(slots) and q % 4 != 3:
                m = int(slots[(q // 2) % len(slots)])
                boxes[b, q] = gt_boxes[b, m] + torch.randn(6, generator=g) * torch.tensor([.15, .15, .15, .1, .1, .1])
                boxes[b, q, 3:] = boxes[b, q, 3:].clamp_min(0.05)
                logits[b, q, int(labels[b, m])] += 3.0
            else:
                boxes[b, q] = torch.cat([torch.randn(3, generator=g) * 3, torch.rand(3, generator=g) * 2 + 0.3])
    gt_boxes = gt_boxes * valid[..., None]
    labels = torch.where(valid, labels, torch.full_like(labels, -1))
    return boxes, logits, {"boxes": gt_boxes, "labels": labels, "valid_mask": valid}


class _Stub(torch.nn.Module):
    def __init__(self, boxes, logits):
        super().__init__()
        self.boxes, self.logits = boxes, logits

    def forward(self, traces, mask):
        lo, hi = int(traces[0, 0, 0]), int(traces[0, 0, 1])
        return {"pred_boxes": self.boxes[lo:hi], "pred_classes": self.logits[lo:hi]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(args.reference, "src", "benchmark"))
    import inference as ref_inf
    import train as ref_train

    out = {}
    for name, (Q, M, B, seed) in CASES.items():
        boxes, logits, targets = case_inputs(name)
        # two batches, to exercise accumulation across batches
        cuts = [(0, B // 2), (B // 2, B)]
        loader = [{"traces": torch.tensor([[[float(lo), float(hi)]]]), "trace_mask": torch.ones(1, 1, dtype=torch.bool),
                   "boxes": targets["boxes"][lo:hi], "labels": targets["labels"][lo:hi], "valid_mask": targets["valid_mask"][lo:hi]}
                  for lo, hi in cuts]
        with contextlib.redirect_stderr(io.StringIO()):
            m = ref_train.evaluate_metrics(_Stub(boxes, logits), loader, torch.device("cpu"), iou_thresh=0.5)
        for k, v in m.items():
            out[f"{name}_metric_{k}"] = np.float64(v)
        keep = -np.ones((B, Q), np.int64)
        for b in range(B):
            with contextlib.redirect_stdout(io.StringIO()):
                preds = ref_inf.post_process_predictions(boxes[b], logits[b], confidence_threshold=0.7, nms_threshold=0.3)
            # identify each emitted prediction by its (unique) box centre
            for i, p in enumerate(preds):
                c = torch.tensor([p["center"]["x"], p["center"]["y"], p["center"]["z"]])
                q = int(torch.argmin((boxes[b, :, :3] - c).abs().sum(1)))
                keep[b, i] = q
                assert LABELS[p["label"]] == int(logits[b, q].argmax())
        out[f"{name}_nms_keep"] = keep
        print(name, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in m.items()}, "kept", (keep >= 0).sum(1).tolist())
    np.savez_compressed(os.path.join(GOLDEN, "eval.npz"), **out)
    print("wrote eval.npz", os.path.getsize(os.path.join(GOLDEN, "eval.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
