"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline may import this; the product path
never does).  numpy restatement of the trace preprocessing of the upstream pipeline:

  * process_traces     <- src/benchmark/inference.py:24-57  (identical body: src/benchmark/dataloader.py:410-457)
  * collate            <- src/benchmark/dataloader.py:510-559 (zero padding to the batch maximum + boolean mask)

PARITY PINNED: tests/golden/features.npz holds outputs of the reference's own process_traces (imported from
/root/reference by oracle/make_golden_features.py) on real dataset traces and edge cases; tests/test_oracle_features.py
checks this restatement against them bit for bit.
"""
from __future__ import annotations

import numpy as np


def process_points(points: np.ndarray, max_len: int = 3000) -> np.ndarray:
    """points: (N, 4) float32 rows (x, y, z, timestamp) in any order -> (min(N, max_len) or 1, 11) float32."""
    pts = np.asarray(points, dtype=np.float32).reshape(-1, 4)
    n = pts.shape[0]
    if n == 0:                                                    # inference.py:34-35
        return np.zeros((1, 11), np.float32)
    pts = pts[np.argsort(pts[:, 3])].copy()                       # :38-39
    pts[:, 3] = pts[:, 3] - pts[0, 3]                             # :40
    step = np.zeros_like(pts)                                     # :43 difference to the previous row, first row 0
    step[1:] = pts[1:] - pts[:-1]
    dt = np.maximum(step[:, 3], np.float32(1e-3))                 # :44
    vel = step[:, :3] / dt[:, None]                               # :45
    acc = np.zeros_like(vel)                                      # :46
    acc[1:] = vel[1:] - vel[:-1]
    sq = vel * vel
    speed = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])             # :47 (norm over 3 components: sequential adds)
    out = np.concatenate([pts, vel, acc, speed[:, None]], axis=1).astype(np.float32)   # :48-50
    if n > max_len:                                               # :53-56
        out = out[downsample_index(n, max_len)]
    return out


def downsample_index(n: int, max_len: int) -> np.ndarray:
    """np.linspace(0, n-1, max_len, dtype=int) spelled out: float64 j*step truncated, last index exactly n-1."""
    step = np.float64(n - 1) / np.float64(max_len - 1)
    idx = (np.arange(max_len, dtype=np.float64) * step).astype(np.int64)
    idx[-1] = n - 1
    return idx


def collate(features: list[np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
    """dataloader.py:531-547: pad each (L_i, 11) with zero rows to max L_i; mask True on real rows."""
    width = max(f.shape[0] for f in features)
    batch = np.zeros((len(features), width, 11), np.float32)
    mask = np.zeros((len(features), width), bool)
    for i, f in enumerate(features):
        batch[i, :f.shape[0]] = f
        mask[i, :f.shape[0]] = True
    return batch, mask
