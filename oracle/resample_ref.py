"""TEST INFRASTRUCTURE ONLY.  numpy definition of how recorded traces become the GRU's fixed-rate input windows
(decision D14; README.md:145 "10 Hz", README.md:147-157 seq_len 500; floor plane (x, z) as in
visualization/plot_data.py:52,73).  Upstream ships no code for this step (PARITY UNPINNED: this file IS the definition;
oracle/make_golden.py used the same lines to cut tests/golden/real_traces.npz)."""
from __future__ import annotations

import numpy as np


def resample_windows(points: np.ndarray, seq_len: int = 500, hz: float = 10.0) -> np.ndarray:
    """points: (N, 4) float64 rows (x, y, z, timestamp) of ONE trace -> (n_windows, seq_len, 2) float32."""
    pts = np.asarray(points, np.float64).reshape(-1, 4)
    if len(pts) < 2:
        return np.zeros((0, seq_len, 2), np.float32)
    pts = pts[np.argsort(pts[:, 3], kind="stable")]
    t = np.arange(pts[0, 3], pts[-1, 3], 1.0 / hz)
    xz = np.stack([np.interp(t, pts[:, 3], pts[:, 0]), np.interp(t, pts[:, 3], pts[:, 2])], axis=-1)
    wins = [xz[s:s + seq_len].astype(np.float32) for s in range(0, len(xz) - seq_len + 1, seq_len)]
    return np.stack(wins) if wins else np.zeros((0, seq_len, 2), np.float32)
