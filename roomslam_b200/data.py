"""Sample-data files in the formats the upstream README specifies (README.md:84-108): one CSV per trace with
`timestamp, x, y` rows and one JSON per trace with the object list.  Host-side plumbing for train.py / evaluate.py."""
from __future__ import annotations

import csv
import glob
import json
import os
from typing import Dict, Tuple

import torch

from . import synth

CLASS_NAMES = ("GROUND", "LOW", "MID", "BLOCK")


def create_sample_data(out_dir: str, n_traces: int = 256, seq_len: int = 500, max_objects: int = 10, seed: int = 0,
                       dt: float = 0.1) -> None:
    os.makedirs(out_dir, exist_ok=True)
    x, tgt = synth.make_sample(n_traces, seq_len, max_objects, seed=seed)
    for i in range(n_traces):
        with open(os.path.join(out_dir, f"trace_{i:05d}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            for t in range(seq_len):
                w.writerow([f"{t * dt:.1f}", f"{x[i, t, 0].item():.6f}", f"{x[i, t, 1].item():.6f}"])
        objs = []
        for k in range(max_objects):
            if tgt["valid"][i, k] > 0:
                c = int(tgt["classes"][i, k])
                objs.append({"class_id": c, "class_name": CLASS_NAMES[c], "position": tgt["positions"][i, k].tolist(),
                             "size": tgt["sizes"][i, k].tolist(), "orientation": float(tgt["orientations"][i, k]),
                             "confidence": 1.0})
        json.dump({"objects": objs}, open(os.path.join(out_dir, f"trace_{i:05d}.json"), "w"))


def load_dir(data_dir: str, seq_len: int = 500, max_objects: int = 10) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    files = sorted(glob.glob(os.path.join(data_dir, "*.csv")))
    if not files:
        raise ValueError(f"no trace CSV files in {data_dir}")        # upstream: ValueError on an empty dataset
    n = len(files)
    x = torch.zeros(n, seq_len, 2)
    tgt = {"classes": torch.zeros(n, max_objects, dtype=torch.long), "positions": torch.zeros(n, max_objects, 2),
           "sizes": torch.ones(n, max_objects, 2), "orientations": torch.zeros(n, max_objects),
           "valid": torch.zeros(n, max_objects)}
    for i, path in enumerate(files):
        rows = [r for r in csv.reader(open(path)) if r]
        pts = torch.tensor([[float(r[1]), float(r[2])] for r in rows[:seq_len]])
        x[i, : len(pts)] = pts
        if len(pts) < seq_len and len(pts) > 0:
            x[i, len(pts):] = pts[-1]
        lab = os.path.splitext(path)[0] + ".json"
        if os.path.exists(lab):
            for k, o in enumerate(json.load(open(lab)).get("objects", [])[:max_objects]):
                tgt["classes"][i, k] = int(o["class_id"])
                tgt["positions"][i, k] = torch.tensor(o["position"])
                tgt["sizes"][i, k] = torch.tensor(o["size"])
                tgt["orientations"][i, k] = float(o.get("orientation", 0.0))
                tgt["valid"][i, k] = 1.0
    return x, tgt


# ---- the shipped benchmark's file formats (src/benchmark/dataloader.py): host-side plumbing for the BiLSTM pipeline ----
COLLIDER_LABELS = {"BLOCK": 0, "LOW": 1, "MID": 2, "HIGH": 3}            # dataloader.py:67-72


def load_trace_points(path: str) -> torch.Tensor:
    """``*_data_*.json``: list of {timestamp, x, y, z} -> (N, 4) fp32 rows (x, y, z, timestamp), the input of
    ``preprocess.trace_features`` (what dataloader.py:420 builds before the kinematic features)."""
    pts = json.load(open(path))
    return torch.tensor([[p["x"], p["y"], p["z"], p["timestamp"]] for p in pts], dtype=torch.float32).reshape(-1, 4)


def colliders_to_targets(colliders, max_colliders: int = 50) -> Dict[str, torch.Tensor]:
    """List of collider dicts ({label, center{x,y,z}, size{x,y,z}}) -> {"boxes" (M, 6), "labels" (M,) int64 (-1 padded),
    "valid_mask" (M,) bool}: the target format of dataloader.py:459-507 (unknown labels map to 0, missing fields to 0.0,
    colliders beyond max_colliders are dropped)."""
    boxes = torch.zeros(max_colliders, 6)
    labels = torch.full((max_colliders,), -1, dtype=torch.long)
    valid = torch.zeros(max_colliders, dtype=torch.bool)
    for i, col in enumerate(colliders[:max_colliders]):
        c, s = col.get("center", {}), col.get("size", {})
        boxes[i] = torch.tensor([c.get("x", 0.0), c.get("y", 0.0), c.get("z", 0.0), s.get("x", 0.0), s.get("y", 0.0), s.get("z", 0.0)])
        labels[i] = COLLIDER_LABELS.get(col.get("label", "BLOCK"), 0)
        valid[i] = True
    return {"boxes": boxes, "labels": labels, "valid_mask": valid}


def load_colliders(path: str, max_colliders: int = 50) -> Dict[str, torch.Tensor]:
    """``colliders.json`` ({"colliders": [...]}, dataset/train/colliders.json:1-19) -> targets of one scene."""
    return colliders_to_targets(json.load(open(path))["colliders"], max_colliders)
