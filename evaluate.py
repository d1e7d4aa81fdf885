#!/usr/bin/env python
"""Thin equivalent of the upstream README's evaluate.py (README.md:76-80; absent upstream).

    python evaluate.py --checkpoint checkpoints/best_model.pth --data_dir data/sample --compare_baseline
Reports the loss terms, class accuracy and axis-aligned IoU over the valid slots (fixed slot assignment, decision D8)
and, with --compare_baseline, the occupancy / stationary summary of the rule-based baseline."""
import argparse
import json

import torch

from roomslam_b200 import OccupancyHeatmapBaseline, RoomSLAM, data


def box_iou(p_pos, p_size, t_pos, t_size):
    lo = torch.maximum(p_pos - p_size / 2, t_pos - t_size / 2)
    hi = torch.minimum(p_pos + p_size / 2, t_pos + t_size / 2)
    inter = (hi - lo).clamp_min(0).prod(-1)
    union = p_size.prod(-1) + t_size.prod(-1) - inter
    return inter / union.clamp_min(1e-9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", default="checkpoints/best_model.pth")
    ap.add_argument("--data_dir", default="data/sample")
    ap.add_argument("--compare_baseline", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args()
    x, tgt = data.load_dir(args.data_dir)
    model = RoomSLAM(precision=args.precision).cuda().eval()
    model.load_state_dict(torch.load(args.checkpoint, map_location="cpu")["model_state_dict"])
    tg = {k: v.cuda() for k, v in tgt.items()}
    with torch.no_grad():
        pred = model(x.cuda())
        loss = model.compute_loss(pred, tg)
    valid = tg["valid"] > 0
    acc = ((pred["class_logits"].argmax(-1) == tg["classes"]) & valid).sum() / valid.sum().clamp_min(1)
    iou = (box_iou(pred["positions"], pred["sizes"], tg["positions"], tg["sizes"]) * valid).sum() / valid.sum().clamp_min(1)
    report = {k: float(v) for k, v in loss.items()}
    report.update(class_accuracy=float(acc), mean_iou=float(iou), n_traces=len(x))
    if args.compare_baseline:
        b = OccupancyHeatmapBaseline()
        occ, stat, dropped = b.bin(x.cuda())
        report["baseline"] = {"visited_cells": int((occ > 0).sum()), "stationary_cells_5s": int(b.stationary_cells(5.0).numel()),
                              "points_outside_room": dropped}
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
