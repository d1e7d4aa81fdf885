"""CPU oracle for the rule-based occupancy-heatmap / stationary-time baseline.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product (roomslam_b200/) never does.

PARITY UNPINNED: upstream names this component (README.md:15 "rule-based occupancy heatmap baseline",
README.md:34 `src/models/baseline.py`, README.md:163-164 "Occupancy heatmaps / Stationary time
analysis") but ships no code, tests or golden vectors for it.  The binning
arithmetic is therefore *defined* here in IEEE fp32, round-to-nearest,
no fused multiply-add, so that a CUDA kernel can reproduce it bit for bit:

  D10  fx = (x - x_min) / res ; ix = floor(fx)            (fp32 sub, fp32 div, floor)
       the point is binned iff 0 <= fx < Gx and 0 <= fy < Gy (NaN/Inf fail both tests); cell = iy*Gx + ix
       Gx = ceil((x_max - x_min)/res - 1e-9) evaluated in fp64 on the host
  D11  sample t >= 1 is stationary iff d2 < thr2 with dx = x_t - x_{t-1}, dy likewise,
       d2 = fl(fl(dx*dx) + fl(dy*dy)), thr2 = float32((v*dt)^2); it is counted in the cell of p_t
       (only if p_t itself is binned).  t = 0 is never stationary.
  D12  one grid aggregated over all traces, int32 counts.

`heatmap_ref.c` in this directory restates the same rules in plain C (used for multi-threaded CPU
timing and for full-size bit-exact checks); tests pin the two against each other and against the
committed golden fixtures.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np


def grid_shape(bounds, resolution) -> Tuple[int, int]:
    x_min, x_max, y_min, y_max = (float(b) for b in bounds)
    gx = int(math.ceil((x_max - x_min) / float(resolution) - 1e-9))
    gy = int(math.ceil((y_max - y_min) / float(resolution) - 1e-9))
    return gy, gx


def bin_points(points: np.ndarray, x_min, y_min, res, gx: int, gy: int, thr2):
    """points (B,T,2) float32 -> (occ int32[gy,gx], stat int32[gy,gx], n_dropped int)."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    assert p.ndim == 3 and p.shape[-1] == 2
    x_min, y_min, res, thr2 = np.float32(x_min), np.float32(y_min), np.float32(res), np.float32(thr2)
    x, y = p[..., 0], p[..., 1]
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        fx = (x - x_min) / res          # fp32 sub then fp32 div (numpy never contracts to FMA)
        fy = (y - y_min) / res
        ok = (fx >= 0) & (fx < np.float32(gx)) & (fy >= 0) & (fy < np.float32(gy))
        ix = np.floor(np.where(ok, fx, 0)).astype(np.int64)
        iy = np.floor(np.where(ok, fy, 0)).astype(np.int64)
        cell = iy * gx + ix
        occ = np.bincount(cell[ok], minlength=gx * gy).astype(np.int32).reshape(gy, gx)
        dx = x[:, 1:] - x[:, :-1]
        dy = y[:, 1:] - y[:, :-1]
        d2 = dx * dx + dy * dy          # two fp32 products, one fp32 add
        st = np.zeros_like(ok)
        st[:, 1:] = (d2 < thr2) & ok[:, 1:]
    stat = np.bincount(cell[st], minlength=gx * gy).astype(np.int32).reshape(gy, gx)
    return occ, stat, int(ok.size - ok.sum())


class OccupancyHeatmapBaseline:
    """Oracle twin of roomslam_b200.OccupancyHeatmapBaseline (same ctor/method signatures)."""

    def __init__(self, bounds=(0.0, 10.0, 0.0, 10.0), resolution: float = 0.05,
                 stationary_speed: float = 0.1, dt: float = 0.1):
        self.bounds = tuple(float(b) for b in bounds)
        self.resolution, self.stationary_speed, self.dt = float(resolution), float(stationary_speed), float(dt)
        self.gy, self.gx = grid_shape(self.bounds, self.resolution)
        self.thr2 = np.float32((self.stationary_speed * self.dt) ** 2)
        self.last_occupancy: Optional[np.ndarray] = None
        self.last_stationary: Optional[np.ndarray] = None
        self.last_dropped = 0

    def bin(self, traces):
        t = np.asarray(traces, dtype=np.float32)
        if t.ndim == 2:
            t = t[None]
        occ, stat, nd = bin_points(t, self.bounds[0], self.bounds[2], self.resolution, self.gx, self.gy, self.thr2)
        self.last_occupancy, self.last_stationary, self.last_dropped = occ, stat, nd
        return occ, stat, nd

    def heatmap(self, traces) -> np.ndarray:
        return self.bin(traces)[0]

    def stationary(self, traces) -> np.ndarray:
        return self.bin(traces)[1]

    def stationary_cells(self, min_seconds: float, traces=None) -> np.ndarray:
        if traces is not None:
            self.bin(traces)
        if self.last_stationary is None:
            raise RuntimeError("stationary_cells() needs traces or a previous stationary()/heatmap() call")
        need = max(1, int(math.ceil(min_seconds / self.dt - 1e-9)))
        return np.flatnonzero(self.last_stationary.reshape(-1) >= need).astype(np.int64)
