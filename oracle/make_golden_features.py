"""Generates tests/golden/features.npz from the REFERENCE's own process_traces (src/benchmark/inference.py:24-57),
imported from /root/reference in the authoring container (it cannot travel to the GPU box; the fixture does).

    python -m oracle.make_golden_features [--reference /root/reference]

Cases: the three shortest real dataset traces (the shortest with a 4000-point cap, the others with the default 3000, so both the
plain and the down-sampled path are pinned), and synthetic edge cases: empty, one point, two points, repeated timestamps
(dt clip), unsorted input, exactly max_len / max_len+1 points with a small cap.  For the long real traces the fixture
keeps every 7th output row plus a SHA-256 of the whole array.
"""
from __future__ import annotations

import argparse
import contextlib
import glob
import hashlib
import io
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_reference(reference: str):
    sys.path.insert(0, os.path.join(reference, "src", "benchmark"))
    if "model" not in sys.modules:
        try:
            import model  # noqa: F401  (inference.py imports build_model at module scope)
        except Exception:  # the model module is irrelevant to process_traces; stub it if its imports are missing
            sys.modules["model"] = types.SimpleNamespace(build_model=None)
    import inference
    return inference.process_traces


def as_dicts(points: np.ndarray):
    return [{"x": float(p[0]), "y": float(p[1]), "z": float(p[2]), "timestamp": float(p[3])} for p in points]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    ref_fn = load_reference(args.reference)

    def ref(points, max_len):
        with contextlib.redirect_stdout(io.StringIO()):
            return ref_fn(as_dicts(points), max_len=max_len).numpy()

    out = {}
    files = sorted(glob.glob(os.path.join(args.reference, "dataset", "*", "*_data_*.json")),
                   key=lambda f: len(json.load(open(f))))[:3]
    for k, f in enumerate(files):
        tr = json.load(open(f))
        pts = np.array([[p["x"], p["y"], p["z"], p["timestamp"]] for p in tr], dtype=np.float64)
        cap = 4000 if k == 0 else 3000                        # the shortest trace (3145 points) also pins the un-capped path
        feats = ref(pts, cap)
        out[f"real{k}_points"] = pts.astype(np.float32)      # float32 of the JSON doubles == what np.array(dtype=f32) makes
        out[f"real{k}_maxlen"] = np.int64(cap)
        out[f"real{k}_rows7"] = feats[::7]
        out[f"real{k}_shape"] = np.array(feats.shape)
        out[f"real{k}_sha256"] = np.frombuffer(hashlib.sha256(feats.tobytes()).digest(), np.uint8)
        print(os.path.basename(f), len(tr), "->", feats.shape)

    rng = np.random.default_rng(20260118)

    def walk(n, unsorted=False, repeats=False):
        t = np.cumsum(rng.uniform(0.0, 0.05, n)) + 100.0
        if repeats:
            t[n // 3:n // 3 + 4] = t[n // 3]
        p = np.stack([np.cumsum(rng.normal(0, 0.02, n)), 1.6 + rng.normal(0, 0.01, n), np.cumsum(rng.normal(0, 0.02, n)), t], 1)
        if unsorted:
            p = p[rng.permutation(n)]
        return p.astype(np.float32).astype(np.float64)

    synthetic = {"empty": (np.zeros((0, 4)), 3000), "one": (walk(1), 3000), "two": (walk(2), 3000),
                 "repeats": (walk(64, repeats=True), 3000), "unsorted": (walk(97, unsorted=True), 3000),
                 "exact_cap": (walk(50), 50), "cap_plus_one": (walk(51), 50), "cap_small": (walk(1000), 17),
                 "cap_two": (walk(33), 2)}
    for name, (pts, cap) in synthetic.items():
        feats = ref(pts, cap)
        out[f"{name}_points"] = pts.astype(np.float32)
        out[f"{name}_maxlen"] = np.int64(cap)
        out[f"{name}_feats"] = feats
        print(name, pts.shape, "->", feats.shape)
    np.savez_compressed(os.path.join(GOLDEN, "features.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "features.npz"), os.path.getsize(os.path.join(GOLDEN, "features.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
