"""Generates tests/golden/colliders.npz: the reference's own TraceColliderDataset._process_colliders
(src/benchmark/dataloader.py:459-507) applied to the real collider file dataset/train/colliders.json, plus the raw
collider list (as a JSON string) so that roomslam_b200.data.colliders_to_targets can be checked against it off-line.

    python -m oracle.make_golden_colliders [--reference /root/reference]
"""
import argparse
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(args.reference, "src", "benchmark"))
    import dataloader as ref_dl
    raw = json.load(open(os.path.join(args.reference, "dataset", "train", "colliders.json")))["colliders"]
    extra = [{"label": "HIGH", "center": {"x": 1.0}, "size": {"y": 2.0}}, {"label": "UNKNOWN"}, {}]     # defaults of .get()
    cases = {"train": raw, "edge": raw[:2] + extra, "overflow": (raw * 6)[:60], "none": []}
    out = {}
    stub = types.SimpleNamespace(max_colliders=50, label_to_id={"BLOCK": 0, "LOW": 1, "MID": 2, "HIGH": 3})
    for name, cols in cases.items():
        boxes, labels, valid = ref_dl.TraceColliderDataset._process_colliders(stub, cols)
        out[f"{name}_json"] = np.array(json.dumps(cols))
        out[f"{name}_boxes"], out[f"{name}_labels"], out[f"{name}_valid"] = boxes.numpy(), labels.numpy(), valid.numpy()
        print(name, len(cols), int(valid.sum()))
    np.savez_compressed(os.path.join(GOLDEN, "colliders.npz"), **out)


if __name__ == "__main__":
    main()
