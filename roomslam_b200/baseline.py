"""Rule-based occupancy-heatmap / stationary-time baseline on B200.

Drop-in for the `src/models/baseline.py` the upstream README names (README.md:15,34,163-164; no code upstream,
SURVEY.md section 0).  Same constructor and method signatures as the CPU oracle (oracle/baseline_ref.py), results
bit-identical; the binning runs in the hand-written sm_100a kernel behind `rs_heatmap_bin` (csrc/heatmap.cu).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib


def grid_shape(bounds, resolution) -> Tuple[int, int]:
    """(gy, gx): ceil(extent/res - 1e-9) in fp64 on the host (decision D10)."""
    x_min, x_max, y_min, y_max = (float(b) for b in bounds)
    gx = int(math.ceil((x_max - x_min) / float(resolution) - 1e-9))
    gy = int(math.ceil((y_max - y_min) / float(resolution) - 1e-9))
    return gy, gx


class OccupancyHeatmapBaseline:
    def __init__(self, bounds=(0.0, 10.0, 0.0, 10.0), resolution: float = 0.05,
                 stationary_speed: float = 0.1, dt: float = 0.1, device: Optional[torch.device] = None):
        self.bounds = tuple(float(b) for b in bounds)
        self.resolution, self.stationary_speed, self.dt = float(resolution), float(stationary_speed), float(dt)
        self.gy, self.gx = grid_shape(self.bounds, self.resolution)
        if self.gx <= 0 or self.gy <= 0:
            raise ValueError(f"empty grid for bounds={bounds} resolution={resolution}")
        self.thr2 = float(torch.tensor((self.stationary_speed * self.dt) ** 2, dtype=torch.float32))
        self.device = torch.device(device) if device is not None else None
        self.last_occupancy: Optional[torch.Tensor] = None
        self.last_stationary: Optional[torch.Tensor] = None
        self.last_dropped = 0
        self._variant = 0

    # -- core ------------------------------------------------------------------------------------------
    def _as_points(self, traces) -> torch.Tensor:
        t = traces if isinstance(traces, torch.Tensor) else torch.as_tensor(traces)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.dim() != 3 or t.shape[-1] != 2:
            raise ValueError(f"traces must have shape (B, T, 2) or (T, 2), got {tuple(t.shape)}")
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        return t

    def bin_into(self, traces: torch.Tensor, occ: torch.Tensor, stat: torch.Tensor, dropped: torch.Tensor,
                 accumulate: bool = False) -> None:
        """Device tensors in, device tensors out, asynchronous on the current stream (the kernel-only path)."""
        if not traces.is_cuda:
            raise _lib.RoomSlamError("bin_into() needs CUDA tensors (no CPU fallback)")
        if not traces.is_contiguous():
            raise ValueError("traces must be contiguous (nothing is re-laid out silently)")
        stream = torch.cuda.current_stream(traces.device).cuda_stream
        with torch.cuda.device(traces.device):
            _lib.call("rs_heatmap_bin_variant", traces.data_ptr(), traces.shape[0], traces.shape[1], self.bounds[0],
                      self.bounds[2], self.resolution, self.gx, self.gy, self.thr2, occ.data_ptr(), stat.data_ptr(),
                      dropped.data_ptr(), int(accumulate), self._variant, stream)

    def bin(self, traces):
        """(occupancy int32 [gy,gx], stationary int32 [gy,gx], n_dropped int).

        CUDA input: kernel on the current stream, outputs on the same device.
        CPU input (numpy / CPU tensor): streamed through `rs_heatmap_bin_host`, outputs are CPU tensors."""
        t = self._as_points(traces)
        if t.is_cuda:
            occ = torch.empty(self.gy, self.gx, dtype=torch.int32, device=t.device)
            stat = torch.empty_like(occ)
            dropped = torch.empty(1, dtype=torch.int64, device=t.device)
            self.bin_into(t.contiguous(), occ, stat, dropped)
            nd = int(dropped.item())
        else:
            t = t.contiguous()
            occ = torch.empty(self.gy, self.gx, dtype=torch.int32)
            stat = torch.empty_like(occ)
            dropped = torch.zeros(1, dtype=torch.int64)
            dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device()) \
                if torch.cuda.is_available() else None
            if dev is None:
                raise _lib.RoomSlamError("no CUDA device: roomslam_b200 has no CPU fallback")
            with torch.cuda.device(dev):
                _lib.call("rs_heatmap_bin_host", t.data_ptr(), t.shape[0], t.shape[1], self.bounds[0], self.bounds[2],
                          self.resolution, self.gx, self.gy, self.thr2, occ.data_ptr(), stat.data_ptr(),
                          dropped.data_ptr())
            nd = int(dropped.item())
        self.last_occupancy, self.last_stationary, self.last_dropped = occ, stat, nd
        return occ, stat, nd

    # -- the README-facing API -----------------------------------------------------------------------------
    def heatmap(self, traces) -> torch.Tensor:
        return self.bin(traces)[0]

    def stationary(self, traces) -> torch.Tensor:
        return self.bin(traces)[1]

    def stationary_cells(self, min_seconds: float, traces=None) -> torch.Tensor:
        if traces is not None:
            self.bin(traces)
        if self.last_stationary is None:
            raise RuntimeError("stationary_cells() needs traces or a previous stationary()/heatmap() call")
        need = max(1, int(math.ceil(min_seconds / self.dt - 1e-9)))
        return torch.nonzero(self.last_stationary.reshape(-1) >= need).reshape(-1).to(torch.int64)

    # -- multi-GPU: shard by trace, exact int32 sum ------------------------------------------
    def bin_distributed(self, local_traces: torch.Tensor, group=None):
        import torch.distributed as dist
        occ, stat, nd = self.bin(local_traces)
        packed = torch.cat([occ.reshape(-1), stat.reshape(-1)])
        nd_t = torch.tensor([nd], dtype=torch.int64, device=packed.device)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(nd_t, op=dist.ReduceOp.SUM, group=group)
        n = self.gx * self.gy
        self.last_occupancy = packed[:n].view(self.gy, self.gx)
        self.last_stationary = packed[n:].view(self.gy, self.gx)
        self.last_dropped = int(nd_t.item())
        return self.last_occupancy, self.last_stationary, self.last_dropped
