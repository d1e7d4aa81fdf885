"""Generates the committed fixtures under tests/golden/ (TEST INFRASTRUCTURE; run in the authoring container).

    python -m oracle.make_golden [--reference /root/reference]

* real_traces.npz   : the 8 real traces of the upstream dataset (dataset/train/*.json, dataset/val/*.json) projected
                      to the floor plane (x, z) (y is height: visualization/plot_data.py:52,73), sorted by timestamp,
                      linearly resampled to 10 Hz (README.md:145) and cut into T=500 windows (decision D14), plus the
                      collider targets projected to 2-D (label map src/benchmark/dataloader.py:67-72 -> D7).
                      Needs /root/reference; everything else below is synthetic.
* heatmap_*.npz     : inputs (or their seeds) and the oracle's occupancy / stationary grids and dropped counts.
* gru_*.npz         : oracle RoomSLAM outputs, losses and gradient digests on seeded inputs and weights.

The upstream repository holds no golden vectors for this path (parity unpinned); these
fixtures pin OUR oracle against accidental drift, and pin the CUDA path to the oracle on the GPU box where
/root/reference does not exist.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import baseline_ref  # noqa: E402
from oracle.room_slam_ref import RoomSLAM  # noqa: E402
from roomslam_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
LABELS = {"GROUND": 0, "LOW": 1, "MID": 2, "BLOCK": 3, "HIGH": 3}  # D7; shipped HIGH has no README class -> BLOCK


def real_traces(reference: str, seq_len: int = 500, hz: float = 10.0):
    windows, names = [], []
    targets = {}
    for split in ("train", "val"):
        col = json.load(open(os.path.join(reference, "dataset", split, "colliders.json")))["colliders"][:10]
        tgt = {"classes": np.zeros(10, np.int64), "positions": np.zeros((10, 2), np.float32),
               "sizes": np.zeros((10, 2), np.float32), "orientations": np.zeros(10, np.float32),
               "valid": np.zeros(10, np.float32)}
        for i, c in enumerate(col):
            tgt["classes"][i] = LABELS.get(c.get("label", "BLOCK"), 3)
            tgt["positions"][i] = (c["center"]["x"], c["center"]["z"])
            tgt["sizes"][i] = (abs(c["size"]["x"]), abs(c["size"]["z"]))
            tgt["valid"][i] = 1.0
        targets[split] = tgt
        for path in sorted(glob.glob(os.path.join(reference, "dataset", split, "*_data_*.json"))):
            pts = json.load(open(path))
            arr = np.array([[p["timestamp"], p["x"], p["z"]] for p in pts], dtype=np.float64)
            arr = arr[np.argsort(arr[:, 0], kind="stable")]
            t = np.arange(arr[0, 0], arr[-1, 0], 1.0 / hz)
            xz = np.stack([np.interp(t, arr[:, 0], arr[:, 1]), np.interp(t, arr[:, 0], arr[:, 2])], axis=-1)
            for s in range(0, len(xz) - seq_len + 1, seq_len):
                windows.append(xz[s:s + seq_len].astype(np.float32))
                names.append(f"{split}/{os.path.basename(path)}@{s}")
    return np.stack(windows), names, targets


def heatmap_edge_points(seed: int = 1) -> np.ndarray:
    """Adversarial binning inputs: exact cell edges k*res in fp32, their neighbours one ulp away, negatives,
    x == x_max, NaN / +-Inf, huge magnitudes, denormals, long exact pauses."""
    rng = np.random.default_rng(seed)
    k = rng.integers(-5, 206, size=(96, 300, 2)).astype(np.float32)
    pts = (k * np.float32(0.05)).astype(np.float32)
    nudge = rng.choice([-1.0, 0.0, 1.0], size=pts.shape).astype(np.float32)
    pts = np.nextafter(pts, pts + nudge).astype(np.float32)
    pts[::7, ::11] = np.nan
    pts[3, 5, 0] = np.inf
    pts[4, 6, 1] = -np.inf
    pts[5, 7] = (1e30, 2.0)
    pts[6, 8] = (-1e30, 2.0)
    pts[7, 9] = (1e-42, 1e-42)
    pts[8, 10] = (10.0, 10.0)
    pts[9, 11] = (np.nextafter(np.float32(10.0), np.float32(0.0)), 0.0)
    pts[10, 12] = (-0.0, -0.0)
    pts[11, 13] = (5e6, 5e6)
    pts[12, 14] = (4194304.0 * 0.05, 1.0)
    pts[20:24, 100:200] = pts[20:24, 99:100]          # long pauses
    pts[24, :, :] = np.float32(3.3)                     # a whole trace in one cell
    pts[25, 1:, :] = pts[25, :-1, :] + np.float32(0.0099)  # slow drift just under the speed threshold
    pts[26, 1:, 0] = pts[26, :-1, 0] + np.float32(0.0101)
    return pts


def make_heatmap(reference_windows: np.ndarray):
    b = baseline_ref.OccupancyHeatmapBaseline()
    out = {}
    edge = heatmap_edge_points()
    occ, stat, nd = b.bin(edge)
    out["edge_points"], out["edge_occ"], out["edge_stat"], out["edge_dropped"] = edge, occ, stat, np.int64(nd)
    for name, (n, t, seed) in {"synth_a": (256, 500, 0), "synth_b": (97, 123, 7), "synth_c": (33, 1, 3),
                               "synth_d": (1, 500, 11), "synth_e": (40, 37, 5)}.items():
        tr = synth.make_traces(n, t, seed=seed).numpy()
        occ, stat, nd = b.bin(tr)
        out[f"{name}_shape"] = np.array([n, t, seed], np.int64)
        out[f"{name}_occ"], out[f"{name}_stat"], out[f"{name}_dropped"] = occ, stat, np.int64(nd)
        out[f"{name}_input_sum"] = np.float64(tr.astype(np.float64).sum())
    # real traces: the upstream room spans x in [-1.6,1.9], z in [-5.9,2.4]
    rb = baseline_ref.OccupancyHeatmapBaseline(bounds=(-2.0, 2.5, -6.5, 3.0), resolution=0.05)
    occ, stat, nd = rb.bin(reference_windows)
    out["real_bounds"] = np.array(rb.bounds, np.float64)
    out["real_occ"], out["real_stat"], out["real_dropped"] = occ, stat, np.int64(nd)
    out["real_cells_5s"] = rb.stationary_cells(5.0)
    np.savez_compressed(os.path.join(GOLDEN, "heatmap.npz"), **out)
    print("heatmap.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "occ" not in k})


def digest(t: torch.Tensor) -> np.ndarray:
    """Small fingerprint of a tensor: [sum, abs-sum, l2, first, last] in float64."""
    d = t.detach().to(torch.float64).reshape(-1)
    return np.array([d.sum().item(), d.abs().sum().item(), d.norm().item(), d[0].item(), d[-1].item()], np.float64)


GRU_CASES = {
    # name: (B, T, H, L, N, seed, dropout-mask?)
    "c1_small": (4, 48, 128, 2, 10, 0, False),
    "c1_mask": (3, 20, 128, 2, 10, 1, True),
    "one_layer": (2, 17, 128, 1, 10, 2, False),
    "h256": (2, 12, 256, 2, 10, 3, False),
    "h64": (5, 9, 64, 2, 6, 4, False),
}


def build_case(name: str):
    B, T, H, L, N, seed, use_mask = GRU_CASES[name]
    torch.manual_seed(1000 + seed)
    model = RoomSLAM(hidden_size=H, num_layers=L, max_objects=N, dropout=0.1 if use_mask else 0.0)
    x, tgt = synth.make_sample(B, T, N, seed=seed)
    mask = None
    if use_mask:
        model.train()
        mask = model.make_dropout_mask(B, T, torch.Generator().manual_seed(seed))
    else:
        model.eval()
    return model, x, tgt, mask


def make_gru():
    out = {}
    for name in GRU_CASES:
        model, x, tgt, mask = build_case(name)
        enc_out, h_n = model.encode(x, mask)
        pred = model(x, mask)
        losses = model.compute_loss(pred, tgt)
        losses["total"].backward()
        out[f"{name}_enc_out"] = enc_out.detach().numpy()
        out[f"{name}_h_n"] = h_n.detach().numpy()
        for k, v in pred.items():
            out[f"{name}_pred_{k}"] = v.detach().numpy()
        out[f"{name}_losses"] = np.array([losses[k].item() for k in
                                          ("total", "class", "position", "size", "orientation", "validity")], np.float64)
        for pn, p in model.named_parameters():
            out[f"{name}_grad_{pn}"] = digest(p.grad)
        out[f"{name}_param_digest"] = np.stack([digest(p) for _, p in model.named_parameters()])
    np.savez_compressed(os.path.join(GOLDEN, "gru.npz"), **out)
    print("gru.npz:", len(out), "arrays")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    windows, names, targets = real_traces(args.reference)
    save = {"windows": windows, "names": np.array(names)}
    for split, tgt in targets.items():
        for k, v in tgt.items():
            save[f"{split}_{k}"] = v
    np.savez_compressed(os.path.join(GOLDEN, "real_traces.npz"), **save)
    print("real_traces.npz:", windows.shape, names[:3])
    make_heatmap(windows)
    make_gru()


if __name__ == "__main__":
    main()
