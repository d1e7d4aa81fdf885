#!/usr/bin/env python
"""Thin equivalent of the upstream README's evaluate.py (README.md:76-80; absent upstream).

    python evaluate.py --checkpoint checkpoints/best_model.pth --data_dir data/sample --compare_baseline
Reports the loss terms, class / validity accuracy, axis-aligned IoU over the valid slots, detection precision / recall and
mAP@0.5 (fixed slot assignment, decision D8; roomslam_b200.evaluation.SlotEvaluator)
and, with --compare_baseline, the occupancy / stationary summary of the rule-based baseline."""
import argparse
import json

import torch

from roomslam_b200 import OccupancyHeatmapBaseline, RoomSLAM, data
from roomslam_b200.evaluation import SlotEvaluator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", default="checkpoints/best_model.pth")
    ap.add_argument("--data_dir", default="data/sample")
    ap.add_argument("--compare_baseline", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "auto"])
    args = ap.parse_args()
    x, tgt = data.load_dir(args.data_dir)
    model = RoomSLAM(precision=args.precision).cuda().eval()
    model.load_state_dict(torch.load(args.checkpoint, map_location="cpu")["model_state_dict"])
    tg = {k: v.cuda() for k, v in tgt.items()}
    with torch.no_grad():
        pred = model(x.cuda())
        loss = model.compute_loss(pred, tg)
    ev = SlotEvaluator(num_classes=pred["class_logits"].shape[-1], iou_thresh=0.5)
    ev.update(pred, tg)
    report = {k: float(v) for k, v in loss.items()}
    report.update(ev.compute(), n_traces=len(x))
    if args.compare_baseline:
        b = OccupancyHeatmapBaseline()
        occ, stat, dropped = b.bin(x.cuda())
        report["baseline"] = {"visited_cells": int((occ > 0).sum()), "stationary_cells_5s": int(b.stationary_cells(5.0).numel()),
                              "points_outside_room": dropped}
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
