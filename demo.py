#!/usr/bin/env python
"""Thin equivalent of the upstream README's demo.py (README.md:58-62; absent upstream): synthetic traces through the
rule-based occupancy-heatmap baseline and an (untrained unless --checkpoint) RoomSLAM model; prints the README's
object-prediction JSON (README.md:93-108)."""
import argparse
import json

import torch

from roomslam_b200 import OccupancyHeatmapBaseline, RoomSLAM, synth
from roomslam_b200.data import CLASS_NAMES


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=64)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    x = synth.make_traces(args.traces, 500, seed=args.seed).cuda()
    b = OccupancyHeatmapBaseline()
    occ, stat, dropped = b.bin(x)
    cells = b.stationary_cells(5.0)
    print(f"heatmap: {int((occ > 0).sum())} visited cells, busiest cell {int(occ.max())} samples, "
          f"{cells.numel()} cells with >= 5 s of standing still, {dropped} samples outside the room")
    model = RoomSLAM().cuda().eval()
    if args.checkpoint:
        model.load_state_dict(torch.load(args.checkpoint, map_location="cpu")["model_state_dict"])
    with torch.no_grad():
        p = model(x[:1])
    conf = torch.sigmoid(p["validity_logits"][0])
    objs = [{"class_id": int(p["class_logits"][0, k].argmax()), "class_name": CLASS_NAMES[int(p["class_logits"][0, k].argmax())],
             "position": [round(v, 3) for v in p["positions"][0, k].tolist()], "size": [round(v, 3) for v in p["sizes"][0, k].tolist()],
             "orientation": round(float(p["orientations"][0, k]), 3), "confidence": round(float(conf[k]), 3)}
            for k in range(conf.numel()) if conf[k] > 0.5]
    print(json.dumps({"objects": objs}, indent=1))


if __name__ == "__main__":
    main()
