"""Generates tests/golden/set_loss.npz from the REFERENCE's own HungarianMatcher + SetCriterion
(src/benchmark/train.py:14-187), imported from /root/reference in the authoring container.

    python -m oracle.make_golden_set_loss [--reference /root/reference]

Cases (seeded): the trained configuration (Q = 30 queries, 50 collider slots) with 0 .. 50 valid colliders per sample
(fewer, equal and more colliders than queries; one sample with none), Q = 80 with few colliders, Q = 5 with many.
Stored: inputs, the matched (pred, gt) index pairs per sample, the four losses and the gradients of total_loss with
respect to pred_boxes / pred_classes.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.set_loss_ref import WEIGHTS  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"trained": (30, 50, 12, 21), "wide": (80, 50, 4, 22), "narrow": (5, 50, 6, 23), "empty": (30, 50, 2, 24)}


def case_inputs(name):
    Q, M, B, seed = CASES[name]
    g = torch.Generator().manual_seed(seed)
    boxes = torch.cat([torch.randn(B, Q, 3, generator=g) * 3, torch.rand(B, Q, 3, generator=g) * 2 + 0.05], -1)
    logits = torch.randn(B, Q, 4, generator=g) * 2
    gt_boxes = torch.cat([torch.randn(B, M, 3, generator=g) * 3, torch.rand(B, M, 3, generator=g) * 2 + 0.05], -1)
    labels = torch.randint(0, 4, (B, M), generator=g)
    counts = torch.randint(1, M + 1, (B,), generator=g)
    if name == "trained":
        counts[:4] = torch.tensor([0, Q, M, 1])
    if name == "empty":
        counts[:] = 0
    if name == "wide":
        counts[:] = torch.tensor([3, 10, 50, 1])
    # valid colliders are not necessarily a prefix of the slots: scatter them
    valid = torch.zeros(B, M, dtype=torch.bool)
    for b in range(B):
        valid[b, torch.randperm(M, generator=g)[: int(counts[b])]] = True
    gt_boxes = gt_boxes * valid[..., None]
    labels = torch.where(valid, labels, torch.full_like(labels, -1))
    return boxes, logits, {"boxes": gt_boxes, "labels": labels, "valid_mask": valid}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(args.reference, "src", "benchmark"))
    import train as ref_train

    out = {}
    for name in CASES:
        boxes, logits, targets = case_inputs(name)
        boxes.requires_grad_(True)
        logits.requires_grad_(True)
        crit = ref_train.SetCriterion(dict(WEIGHTS))
        losses = crit({"pred_boxes": boxes, "pred_classes": logits}, targets)
        pairs = crit.matcher.forward(boxes.detach(), logits.detach(), targets["boxes"], targets["labels"], targets["valid_mask"])
        if losses["total_loss"].requires_grad:
            losses["total_loss"].backward()
        for k in ("class_loss", "l1_loss", "giou_loss", "total_loss"):
            out[f"{name}_{k}"] = np.float32(float(losses[k]))
        out[f"{name}_dboxes"] = (boxes.grad if boxes.grad is not None else torch.zeros_like(boxes)).numpy()
        out[f"{name}_dlogits"] = (logits.grad if logits.grad is not None else torch.zeros_like(logits)).numpy()
        K = min(CASES[name][0], CASES[name][1])
        pi = -np.ones((len(pairs), K), np.int64)
        gi = -np.ones((len(pairs), K), np.int64)
        for b, (p, q) in enumerate(pairs):
            pi[b, : len(p)] = np.asarray(p, np.int64)
            gi[b, : len(q)] = np.asarray(q, np.int64)
        out[f"{name}_pred_idx"], out[f"{name}_gt_idx"] = pi, gi
        print(name, {k: round(float(v), 5) for k, v in losses.items()}, "pairs", [len(p) for p, _ in pairs])
    np.savez_compressed(os.path.join(GOLDEN, "set_loss.npz"), **out)
    print("wrote set_loss.npz", os.path.getsize(os.path.join(GOLDEN, "set_loss.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
