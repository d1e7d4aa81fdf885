"""On-GPU trace preprocessing: raw (x, y, z, timestamp) points -> padded (B, W, 11) kinematic features + mask.

Mirrors, for a whole batch in one launch, what the upstream pipeline does per item on the host in numpy:
``process_traces`` (src/benchmark/inference.py:24-57, same body in src/benchmark/dataloader.py:410-457) followed by the
padding of ``collate_fn`` (src/benchmark/dataloader.py:510-559).  Bit-identical to the numpy code (rs_trace_features).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib

FEATURES = ("x", "y", "z", "t", "vx", "vy", "vz", "ax", "ay", "az", "speed")


def pack_traces(traces: Sequence, device="cuda") -> tuple[torch.Tensor, torch.Tensor]:
    """List of (N_i, 4) arrays / tensors (rows x, y, z, timestamp) -> packed (total, 4) fp32 on `device`, offsets (B+1) int64 (CPU)."""
    parts, offs = [], [0]
    for t in traces:
        t = torch.as_tensor(np.asarray(t, dtype=np.float32) if not torch.is_tensor(t) else t, dtype=torch.float32).reshape(-1, 4)
        parts.append(t)
        offs.append(offs[-1] + t.shape[0])
    packed = torch.cat(parts) if parts else torch.zeros(0, 4)
    return packed.to(device, non_blocking=True).contiguous(), torch.tensor(offs, dtype=torch.int64)


def sort_by_time(packed: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """Orders every trace by timestamp (inference.py:38-39).  Ties keep their input order (numpy's argsort leaves them
    unspecified)."""
    if packed.shape[0] == 0:
        return packed
    counts = (offsets[1:] - offsets[:-1]).to(packed.device)
    seg = torch.repeat_interleave(torch.arange(counts.numel(), device=packed.device), counts)
    by_t = torch.sort(packed[:, 3], stable=True).indices
    order = by_t[torch.sort(seg[by_t], stable=True).indices]
    return packed[order].contiguous()


@_lib.on_tensor_device
def trace_features(traces, offsets: torch.Tensor | None = None, max_len: int = 3000, sort="auto",
                   check_sorted: bool = False) -> dict:
    """traces: list of (N_i, 4) arrays, or a packed (total, 4) CUDA tensor with `offsets` (B+1 int64).

    sort: "auto" (default) runs the kernel on the input order and re-runs it on time-sorted points only if the kernel saw
    a decreasing timestamp (recorded traces are already ordered; the two device sorts cost ~40x the kernel);
    True always sorts (the reference's unconditional argsort, inference.py:38-39); False never does.

    Returns {"traces": (B, W, 11) fp32, "trace_mask": (B, W) bool, "lengths": (B,) int64}, W = max_i min(N_i, max_len)
    (an empty trace gives one zero row, like the reference) -- the keys collate_fn emits (dataloader.py:549-551).
    """
    if offsets is None:
        packed, offsets = pack_traces(traces)
    else:
        packed = traces
    if not packed.is_cuda:
        raise _lib.RoomSlamError("trace_features: the points must be on a CUDA device (there is no CPU path)")
    if max_len < 2:
        raise ValueError("max_len must be >= 2")
    packed = packed.contiguous()
    host_off = offsets.cpu()
    B = host_off.numel() - 1
    counts = host_off[1:] - host_off[:-1]
    if B > 0 and (int(counts.min()) < 0 or int(host_off[-1]) != packed.shape[0] or int(host_off[0]) != 0):
        raise ValueError("offsets must start at 0, be non-decreasing and end at the number of points")
    width = int(torch.clamp(counts, 1, max_len).max()) if B > 0 else 1
    if sort is True:
        packed = sort_by_time(packed, host_off)
    dev = packed.device
    feats = torch.empty(B, width, 11, dtype=torch.float32, device=dev)
    mask = torch.empty(B, width, dtype=torch.uint8, device=dev)
    lengths = torch.empty(B, dtype=torch.int64, device=dev)
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    dev_off = host_off.to(dev, non_blocking=True)
    _lib.call("rs_trace_features", packed.data_ptr(), dev_off.data_ptr(), B, max_len, width, feats.data_ptr(),
              mask.data_ptr(), lengths.data_ptr(), flag.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if sort == "auto" and int(flag.item()):
        return trace_features(sort_by_time(packed, host_off), host_off, max_len, sort=False, check_sorted=check_sorted)
    if check_sorted and int(flag.item()):
        raise ValueError("trace_features: a trace is not sorted by timestamp (pass sort=True)")
    return {"traces": feats, "trace_mask": mask.bool(), "lengths": lengths}


def resample_windows(traces: Sequence, seq_len: int = 500, hz: float = 10.0):
    """Recorded traces -> the GRU model's input: list of (N_i, 4) arrays with rows (x, y, z, timestamp) (float64 keeps
    the JSON precision) -> {"windows": (W, seq_len, 2) fp32 on the GPU (floor plane x, z, resampled to `hz` by linear
    interpolation, cut into non-overlapping windows), "trace": (W,) index of the source trace}.  Bit-identical to
    numpy.arange + numpy.interp (rs_resample_windows_f64)."""
    parts, offs, win_trace, win_start = [], [0], [], []
    step = 1.0 / hz
    for b, t in enumerate(traces):
        t = torch.as_tensor(np.asarray(t, dtype=np.float64) if not torch.is_tensor(t) else t, dtype=torch.float64).reshape(-1, 4).cpu()
        if t.shape[0] >= 2:
            t = t[torch.sort(t[:, 3], stable=True).indices]
            n = int(np.ceil((float(t[-1, 3]) - float(t[0, 3])) / step))          # len(numpy.arange(t0, t_last, step))
            for s in range(0, n - seq_len + 1, seq_len):
                win_trace.append(b)
                win_start.append(s)
        parts.append(t)
        offs.append(offs[-1] + t.shape[0])
    W = len(win_trace)
    out = torch.empty(W, seq_len, 2, dtype=torch.float32, device="cuda")
    if W:
        packed = torch.cat(parts).cuda().contiguous()
        dev = lambda v: torch.tensor(v, dtype=torch.int64, device="cuda")      # noqa: E731
        o, wt, ws = dev(offs), dev(win_trace), dev(win_start)
        _lib.call("rs_resample_windows_f64", packed.data_ptr(), o.data_ptr(), wt.data_ptr(), ws.data_ptr(), W, seq_len, step,
                  out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return {"windows": out, "trace": torch.tensor(win_trace, dtype=torch.int64)}
